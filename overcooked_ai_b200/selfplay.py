"""Self-play rollout collection with a torch policy on top of the batched engine (BASELINE config 5).

The reference collects PPO rollouts with Ray workers that each run one Python env and a TF copy of
the policy (human_aware_rl/rllib/rllib.py:293-342, ppo/ppo_rllib.py:7-80).  Here one process per GPU
keeps N environments on the device and runs, per transition,

    lossless_state_encoding (K2, fp32 [N,2,W,H,26])  ->  policy CNN (torch)  ->  multinomial
    ->  ovc_step (K1)  ->  reward accumulation

with no host round trip; the whole transition can be captured in one CUDA graph.  The observation
tensor is consumed zero-copy: ``[N,2,W,H,26]`` viewed as ``(2N, 26, W, H)`` is exactly torch's
channels-last memory format.

The policy is shaped like the reference's ``RllibPPOModel`` defaults (ppo_rllib.py:43-79 with
ppo_rllib_client.py:85-88: conv 5x5x25 'same', conv 3x3x25 'same', conv 3x3x25 'valid', 3 dense layers
of 64, leaky ReLU, heads 6 + 1), random init, shared by both agents.  It is a CONSUMER of the hot
path (library kernels: cuDNN / cuBLAS through torch), not part of it.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class RllibShapedCNN(nn.Module):
    def __init__(self, width, height, in_planes=26, num_filters=25, hidden=64, num_hidden_layers=3, num_actions=6):
        super().__init__()
        self.conv_initial = nn.Conv2d(in_planes, num_filters, 5, padding=2)
        self.conv_0 = nn.Conv2d(num_filters, num_filters, 3, padding=1)
        self.conv_1 = nn.Conv2d(num_filters, num_filters, 3, padding=0)
        flat = num_filters * (width - 2) * (height - 2)
        dims = [flat] + [hidden] * num_hidden_layers
        self.dense = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_hidden_layers)])
        self.logits = nn.Linear(hidden, num_actions)
        self.value = nn.Linear(hidden, 1)

    def forward(self, obs_nchw):
        x = F.leaky_relu(self.conv_initial(obs_nchw), 0.2)
        x = F.leaky_relu(self.conv_0(x), 0.2)
        x = F.leaky_relu(self.conv_1(x), 0.2)
        x = x.flatten(1)
        for d in self.dense:
            x = F.leaky_relu(d(x), 0.3)
        return self.logits(x), self.value(x).squeeze(-1)


class DenseGridPolicy(nn.Module):
    """The same network as ``RllibShapedCNN`` with every convolution folded into ONE matrix per layer.

    On a 5x4 (or 9x5) grid a 'same' convolution spends most of its taps on padding: conv 5x5x26->25 over 20 cells is
    325 k MACs per observation, while the linear map it IS — 520 inputs -> 500 outputs — is 260 k.  cuDNN also runs
    25/26-channel convolutions far below the tensor-core peak, whereas ``[2N, 520] x [520, 500]`` is a plain library
    GEMM.  The matrices are built once by pushing the identity through each convolution (exact: same weights, same
    function, only the summation order differs), in the observation kernel's own element order ``[x][y][channel]``, so
    K2's output is consumed as ``[2N, W*H*26]`` without any permute.  The policy stays a CONSUMER of the hot path
    (library GEMMs), not part of it."""

    def __init__(self, cnn, width, height):
        super().__init__()
        self.W, self.H = width, height
        with torch.no_grad():
            mats = []
            shape = (26, width, height)  # (channels, x, y) as the conv sees it
            for conv in (cnn.conv_initial, cnn.conv_0, cnn.conv_1):
                c, w, h = shape
                n_in = c * w * h
                # basis vector k of the flat [x][y][c] input -> NCHW image with a single one
                eye = torch.eye(n_in, dtype=conv.weight.dtype, device=conv.weight.device).view(n_in, w, h, c).permute(0, 3, 1, 2)
                out = F.conv2d(eye, conv.weight, None, padding=conv.padding)  # (n_in, c_out, w', h'), bias added separately
                co, wo, ho = out.shape[1:]
                mats.append((out.permute(0, 2, 3, 1).reshape(n_in, wo * ho * co).t().contiguous(),   # [n_out, n_in], [x][y][c] order
                             conv.bias.view(1, 1, co).expand(wo, ho, co).reshape(-1).clone()))
                shape = (co, wo, ho)
            self.conv_as_linear = nn.ModuleList()
            for m, b in mats:
                lin = nn.Linear(m.shape[1], m.shape[0])
                lin.weight.copy_(m), lin.bias.copy_(b)
                self.conv_as_linear.append(lin)
            # the first dense layer consumed conv_1's NCHW flatten (c, x, y): re-order its inputs to [x][y][c]
            co, wo, ho = shape
            first = cnn.dense[0]
            d0 = nn.Linear(first.in_features, first.out_features)
            d0.weight.copy_(first.weight.view(-1, co, wo, ho).permute(0, 2, 3, 1).reshape(first.out_features, -1))
            d0.bias.copy_(first.bias)
            self.dense = nn.ModuleList([d0] + [_clone_linear(d) for d in cnn.dense[1:]])
            self.logits, self.value = _clone_linear(cnn.logits), _clone_linear(cnn.value)

    def forward(self, obs_flat):
        """obs_flat: [2N, W*H*26] in K2's element order."""
        x = obs_flat
        for lin in self.conv_as_linear:
            x = F.leaky_relu(lin(x), 0.2)
        for d in self.dense:
            x = F.leaky_relu(d(x), 0.3)
        return self.logits(x), self.value(x).squeeze(-1)


def _clone_linear(l):
    c = nn.Linear(l.in_features, l.out_features)
    with torch.no_grad():
        c.weight.copy_(l.weight), c.bias.copy_(l.bias)
    return c


class SelfPlayRollout(object):
    """Policy-in-the-loop rollout: both agents of every environment act from the same network."""

    def __init__(self, env, model=None, autocast_dtype=torch.bfloat16, use_graph=True, reward_shaping_factor=1.0,
                 obs_dtype=None, dense=True):
        """obs_dtype: element type K2 writes (default: bfloat16 when the policy runs under bf16 autocast — the plane
        values are exact in bf16 and the conversion pass disappears — else float32).
        dense: evaluate the network through ``DenseGridPolicy`` (one library GEMM per layer) instead of cuDNN convolutions."""
        assert len({(l.width, l.height) for l in env.layouts}) == 1, "one grid shape per rollout (group envs by layout)"
        self.env = env
        l = env.layouts[0]
        self.W, self.H = l.width, l.height
        dev = env.device
        self.model = (model or RllibShapedCNN(self.W, self.H)).to(dev).to(memory_format=torch.channels_last).eval()
        self.dense_model = DenseGridPolicy(self.model, self.W, self.H).to(dev).eval() if dense else None
        self.autocast_dtype = autocast_dtype
        self.factor = float(reward_shaping_factor)
        N = env.n_envs
        if obs_dtype is None:
            obs_dtype = torch.bfloat16 if autocast_dtype == torch.bfloat16 else torch.float32
        self.obs = torch.empty((N, 2, self.W, self.H, 26), dtype=obs_dtype, device=dev)
        self.actions = torch.zeros((N, 2), dtype=torch.int32, device=dev)
        self.ret_sparse = torch.zeros(N, dtype=torch.int64, device=dev)      # running episode return (sparse)
        self.ret_mixed = torch.zeros(N, dtype=torch.float32, device=dev)    # sparse + factor * shaped (rllib.py:328-329)
        self.values = torch.zeros((N, 2), dtype=torch.float32, device=dev)
        self.graph = None
        self.use_graph = use_graph

    def _transition(self):
        env = self.env
        env.lossless_state_encoding(out=self.obs)  # K2
        with torch.no_grad(), torch.autocast("cuda", dtype=self.autocast_dtype, enabled=self.autocast_dtype is not None):
            if self.dense_model is not None:
                logits, value = self.dense_model(self.obs.view(2 * env.n_envs, self.W * self.H * 26))
            else:
                x = self.obs.view(2 * env.n_envs, self.W, self.H, 26).permute(0, 3, 1, 2)  # (2N,26,W,H), channels-last strides
                logits, value = self.model(x)
        probs = torch.softmax(logits.float(), dim=-1)
        a = torch.multinomial(probs, 1).view(env.n_envs, 2)
        self.actions.copy_(a)
        self.values.copy_(value.float().view(env.n_envs, 2))
        sparse, shaped, done, events = env.step(self.actions)  # K1 (auto-reset inside)
        self.ret_sparse += sparse
        self.ret_mixed += sparse.float() + self.factor * shaped.sum(-1).float()

    def run(self, n_steps):
        """Advance every environment n_steps transitions; returns the number of env-steps done."""
        if self.use_graph and self.graph is None:
            # warm-up + capture must not advance the environments: snapshot, then restore
            saved = (self.env.state.clone(), self.ret_sparse.clone(), self.ret_mixed.clone())
            s = torch.cuda.Stream(self.env.device)
            s.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(s):
                for _ in range(3):
                    self._transition()
            torch.cuda.current_stream(self.env.device).wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._transition()
            self.env.state.copy_(saved[0]), self.ret_sparse.copy_(saved[1]), self.ret_mixed.copy_(saved[2])
        for _ in range(n_steps):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._transition()
        return n_steps * self.env.n_envs

    def env_only(self, n_steps):
        """The same transitions without the policy: encode + step with the last sampled actions
        (used to report the env-only share of the pipeline)."""
        for _ in range(n_steps):
            self.env.lossless_state_encoding(out=self.obs)
            self.env.step(self.actions)
        return n_steps * self.env.n_envs
