"""Drop-in single-environment adapters with the reference's call surface.

``OvercookedGridworld`` keeps the reference's signatures for the hot path
(src/overcooked_ai_py/mdp/overcooked_mdp.py): ``from_layout_name`` :1151, ``from_grid`` :1175,
``get_standard_start_state`` :1297, ``get_state_transition`` :1375-1430,
``lossless_state_encoding`` :2385-2561, ``featurize_state`` :2579-2898 — and runs each of them as an
N=1 launch of the same CUDA kernels the batched engine uses: pack -> launch -> unpack.  It exists
for API compatibility and parity testing; one launch + sync per call cannot beat the reference's
~40 us Python step, throughput comes from ``BatchedOvercookedEnv`` (SURVEY.md §7 "hard parts").

Error behaviour follows the reference: ``ValueError`` for an illegal action (:1394-1398),
``AssertionError`` for an invalid state (_check_valid_state :1910-1949).
"""
import itertools
from collections import defaultdict

import numpy as np
import torch

from overcooked_ai_b200 import layout as L
from overcooked_ai_b200.actions import Action
from overcooked_ai_b200.batched import BatchedOvercookedEnv
from overcooked_ai_b200.layout import EVENT_TYPES
from overcooked_ai_b200.state import ObjectState, OvercookedState, SoupState


class OvercookedGridworld(object):
    def __init__(self, compiled, device="cuda"):
        assert isinstance(compiled, L.CompiledLayout)
        self.compiled = compiled
        self.device = device
        self._engine = None
        # attributes the reference exposes (:1128-1148)
        c = compiled
        self.layout_name = c.layout_name
        self.height, self.width, self.shape = c.height, c.width, (c.width, c.height)
        self.terrain_mtx = c.terrain_mtx
        self.terrain_pos_dict = c.terrain_pos_dict
        self.start_player_positions = c.start_player_positions
        self.num_players = c.num_players
        self.start_bonus_orders = c.start_bonus_orders
        self.start_all_orders = c.start_all_orders
        self.reward_shaping_params = c.reward_shaping_params
        self.order_bonus = c.order_bonus
        self.old_dynamics = c.old_dynamics
        self.recipe_config = c.recipe_config

    # ---- construction ----------------------------------------------------------------------------
    @staticmethod
    def from_layout_name(layout_name, **params_to_overwrite):
        """:1151-1172.  ``device=...`` (not a layout parameter) selects the CUDA device of the N = 1 engine."""
        device = params_to_overwrite.pop("device", "cuda")
        return OvercookedGridworld(L.compile_layout(layout_name, **params_to_overwrite), device=device)

    @staticmethod
    def from_grid(layout_grid, base_layout_params={}, params_to_overwrite={}, debug=False, device="cuda"):
        params = dict(base_layout_params)
        params.update(params_to_overwrite)
        name = params.pop("layout_name", "|".join("".join(r) for r in layout_grid))
        return OvercookedGridworld(L.CompiledLayout(name, [str(r) if isinstance(r, str) else "".join(r) for r in layout_grid], **params), device=device)

    # ---- layout queries (same names as the reference, :1733-1807) --------------------------------
    def get_valid_player_positions(self):
        return self.terrain_pos_dict[" "]

    def get_terrain_type_at_pos(self, pos):
        return self.compiled.get_terrain_type_at_pos(pos)

    def get_pot_locations(self):
        return list(self.terrain_pos_dict["P"])

    def get_counter_locations(self):
        return list(self.terrain_pos_dict["X"])

    def get_onion_dispenser_locations(self):
        return list(self.terrain_pos_dict["O"])

    def get_tomato_dispenser_locations(self):
        return list(self.terrain_pos_dict["T"])

    def get_dish_dispenser_locations(self):
        return list(self.terrain_pos_dict["D"])

    def get_serving_locations(self):
        return list(self.terrain_pos_dict["S"])

    @property
    def num_pots(self):
        return self.compiled.n_pots

    def get_standard_start_state(self):
        return self.compiled.get_standard_start_state()

    def get_valid_joint_player_positions(self):
        """All ordered tuples of distinct floor cells, in itertools.product order (:1736-1747)."""
        cells = self.get_valid_player_positions()
        return [j for j in itertools.product(cells, repeat=self.num_players) if len(set(j)) == len(j)]

    def get_pot_states(self, state):
        """{'empty' | '<k>_items' | 'cooking' | 'ready': [pot positions]} (:1809-1838)."""
        out = defaultdict(list)
        for pos in self.get_pot_locations():
            if not state.has_object(pos):
                out["empty"].append(pos)
                continue
            soup = state.get_object(pos)
            assert soup.name == "soup", "soup at %s is not a soup but a %s" % (pos, soup.name)
            tick, ct = soup._cooking_tick, self.compiled.soup_cook_time(soup) if soup._cooking_tick >= 0 else None
            if tick >= 0 and tick >= ct:
                out["ready"].append(pos)
            elif tick >= 0:
                out["cooking"].append(pos)
            else:
                out["%d_items" % len(soup.ingredients)].append(pos)
        return out

    def get_actions(self, state):
        """Every action is legal for every player in every (valid) state (:1273-1288)."""
        L.pack_state(self.compiled, state, 0, self.compiled.state_words)  # _check_valid_state: AssertionError if invalid
        return [list(Action.ALL_ACTIONS) for _ in state.players]

    def get_random_start_state_fn(self, random_start_pos=False, rnd_obj_prob_thresh=0.0):
        """The host form of the reference's randomised start states (:1307-1369), for the drop-in
        ``OvercookedEnv(start_state_fn=...)``.  numpy's GLOBAL generator is consumed exactly as the reference consumes
        it — one ``choice`` for the joint position; per pot ``rand``, and if that fills it ``randint`` x2 + ``rand``;
        per player ``rand``, and if that hands it something ``choice(p=[.2, .6, .2])`` + ``randint`` x2 (drawn whatever
        the object is) — so a seeded run starts from the reference's states.  (The batched engine draws on the
        device instead: ovc_random_start_t.)"""
        layout, thr = self.compiled, rnd_obj_prob_thresh

        def draw_counts():
            onions = int(np.random.randint(low=1, high=4))
            return onions, int(np.random.randint(low=0, high=4 - onions))

        def cook_time_of(onions, tomatoes):
            return int(layout.cook_time[onions * 4 + tomatoes])

        def make():
            cells = self.start_player_positions
            if random_start_pos:
                joint = self.get_valid_joint_player_positions()
                cells = joint[np.random.choice(len(joint))]
            state = OvercookedState.from_player_positions(cells, bonus_orders=self.start_bonus_orders,
                                                          all_orders=self.start_all_orders)
            if thr == 0:
                return state
            for pot in self.get_pot_locations():  # a fresh state has no soups: every pot is "empty" (:1331)
                if np.random.rand() < thr:
                    o, t = draw_counts()
                    tick = 0 if np.random.rand() < thr else -1  # cooking from tick 0, or idle
                    state.add_object(SoupState.get_soup(pot, num_onions=o, num_tomatoes=t, cooking_tick=tick,
                                                        cook_time=cook_time_of(o, t)))
            for player in state.players:
                if np.random.rand() < thr:
                    kind = str(np.random.choice(["dish", "onion", "soup"], p=[0.2, 0.6, 0.2]))
                    o, t = draw_counts()
                    held = ObjectState(kind, player.position) if kind != "soup" else SoupState.get_soup(
                        player.position, num_onions=o, num_tomatoes=t, finished=True, cook_time=cook_time_of(o, t))
                    player.set_object(held)
            return state

        return make

    def soup_cook_time(self, soup):
        return self.compiled.soup_cook_time(soup)

    def is_terminal(self, state):
        return False

    def get_lossless_state_encoding_shape(self):
        return np.array(list(self.shape) + [26])

    def get_featurize_state_shape(self, num_pots=2):
        return (self.num_players * (num_pots * 10 + 28),)

    # ---- the N=1 engine ----------------------------------------------------------------------------
    def _eng(self):
        if self._engine is None:
            self._engine = BatchedOvercookedEnv(self.compiled, 1, horizon=0, device=self.device)
            self._act = torch.zeros((1, 2), dtype=torch.int32, device=self._engine.device)
        return self._engine

    def _load(self, state):
        eng = self._eng()
        rec = L.pack_state(self.compiled, state, 0, eng.state_words)  # AssertionError on invalid states
        eng.state.copy_(torch.from_numpy(rec).view(1, -1))
        return eng

    def get_state_transition(self, state, joint_action, display_phi=False, motion_planner=None):
        """(new_state, infos) exactly like the reference (:1375-1430); infos has ``event_infos``
        (25 event names -> [bool, bool]), ``sparse_reward_by_agent`` and ``shaped_reward_by_agent``."""
        if len(joint_action) != 2:
            raise ValueError("Illegal action %s in state %s" % (joint_action, state))
        try:
            idx = [Action.to_index(a) for a in joint_action]
        except ValueError:
            raise ValueError("Illegal action %s in state %s" % (joint_action, state))
        eng = self._load(state)
        phi_s = float(eng.potential(0.99)[0].item()) if display_phi else None
        self._act.copy_(torch.tensor([idx], dtype=torch.int32))
        sparse, shaped, done, events = eng.step(self._act)
        rec = eng.state[0].cpu().numpy()
        ev = events[0].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        shaped = shaped[0].cpu().numpy()
        new_state = L.unpack_state(self.compiled, rec)
        events_infos = {name: [bool((int(ev[a]) >> i) & 1) for a in range(2)] for i, name in enumerate(EVENT_TYPES)}
        per_agent = [
            int(self.compiled.deliver_value[(int(ev[a]) >> L.EV_RECIPE_SHIFT) & 15]) * ((int(ev[a]) >> 15) & 1)
            for a in range(2)
        ]
        assert sum(per_agent) == int(sparse[0].item())
        infos = {
            "event_infos": events_infos,
            "sparse_reward_by_agent": per_agent,
            "shaped_reward_by_agent": [int(shaped[0]), int(shaped[1])],
        }
        if display_phi:  # :1422-1429
            infos["phi_s"] = phi_s
            infos["phi_s_prime"] = float(eng.potential(0.99)[0].item())
        return new_state, infos

    def potential_function(self, state, mp=None, gamma=0.99):
        """phi(state) (:2920-3250).  ``mp`` is accepted for signature compatibility and ignored: the planner
        costs of the default NO_COUNTERS_PARAMS planner are baked into the layout's cost table."""
        return float(self._load(state).potential(gamma)[0].item())

    def lossless_state_encoding(self, overcooked_state, horizon=400, debug=False):
        """Tuple of two (W, H, 26) int64 arrays, one per player (:2385-2561)."""
        eng = self._load(overcooked_state)
        saved = eng.horizon
        eng.horizon = int(horizon) if horizon < 2**31 else 0
        try:
            enc = eng.lossless_state_encoding(dtype=torch.int32)[0].cpu().numpy().astype(np.int64)
        finally:
            eng.horizon = saved
        return (enc[0], enc[1])

    def featurize_state(self, overcooked_state, mlam=None, num_pots=2, **kwargs):
        """List of two float64 vectors (:2579-2898).  ``mlam`` is accepted for signature
        compatibility and ignored: the planner distances (default NO_COUNTERS_PARAMS) are baked into
        the layout's lookup table (layout.CompiledLayout.feature_lut)."""
        eng = self._load(overcooked_state)
        f = eng.featurize_state(num_pots=num_pots)[0].cpu().numpy().astype(np.float64)
        return [f[0], f[1]]

    @property
    def mdp_params(self):
        return {
            "layout_name": self.layout_name,
            "terrain": self.terrain_mtx,
            "start_player_positions": self.start_player_positions,
            "start_bonus_orders": self.start_bonus_orders,
            "rew_shaping_params": dict(self.reward_shaping_params),
            "start_all_orders": self.start_all_orders,
        }


__all__ = ["OvercookedGridworld", "OvercookedState", "EVENT_TYPES"]
