// ovc_wide.cuh — K9 wide_layers_kernel (included by ovc_b200.cu after ovc_tail.cuh): the two wide layers of the rollout
// policy between K7 and K8 (reference model: human_aware_rl/ppo/ppo_rllib.py:54-62, the two 3x3 convolutions, each folded
// into one matrix by the host) as ONE tcgen05 kernel:
//
//     a1 = leaky_relu(a0 . W1^T + b1)        a0 [M][512] bf16 (K7's output), W1 [512][512] bf16
//     z2 = a1 . W2^T + b2                    W2 [160][512] bf16, z2 [M][160] bf16 (K8's input: its leaky ReLU is applied there)
//
// The 128 x 512 activation tile a1 never reaches HBM: layer 1 accumulates in TMEM (all 512 columns: two N = 256
// tcgen05.mma per k-step), the epilogue warps read it back (tcgen05.ld), add the bias, apply the leaky ReLU, round to
// bf16 and write it to shared memory in the K-major 128-byte-swizzle layout a UMMA descriptor reads — the A operand of
// layer 2, whose accumulator re-uses TMEM columns 0..159.  As library calls the same work is two GEMMs and an activation
// pass with a1 written once and read twice (3 x 67 MB per 65 536 rows).
//
// Persistent: one CTA per SM walks 128-row tiles; 10 warps: warp 0 = TMA producer (one lane), warp 1 = TMEM allocation + MMA issue (one lane),
// warps 2-9 = epilogue (TMEM lane quadrant = warp & 3, two warps per quadrant share the columns).  Operand tiles arrive by
// TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) through mbarrier rings.  Layer 1 is computed one N-half at a time
// (columns 0-255, then 256-511: 8 k-chunks of {a0 128 x 64, W1 256 x 64} = 48 KB each through three stages), so that the
// epilogue of the first half — TMEM -> registers -> the first four k-chunks of the a1 tile — runs UNDER the second half's
// MMAs; layer 2's first four k-steps then run under the second half's epilogue.  Shared memory: a1 chunks 0-3 (64 KB) own
// their space, the three stages (144 KB) follow; once layer 1 is done a1 chunks 4-7 and W2's two-slot ring (160 x 64,
// 20 KB per chunk) overlay the stages.  Every mbarrier wait is bounded (~10 s; a protocol error traps instead of hanging the GPU).
#pragma once
#include <cuda_bf16.h>

namespace ovc {

constexpr int WL_BM = 128, WL_BK = 64, WL_K0 = 512, WL_N1 = 512, WL_N2 = 160;
constexpr int WL_KC = WL_K0 / WL_BK;        // k-chunks of layer 1
constexpr int WL_KC2 = WL_N1 / WL_BK;       // k-chunks of layer 2
constexpr int WL_THREADS = 320;
constexpr int WL_STAGES = 3;
constexpr uint32_t WL_A_BYTES = WL_BM * WL_BK * 2;            // 16 KB
constexpr uint32_t WL_B1_BYTES = 256 * WL_BK * 2;             // 32 KB: one N-half of a W1 k-chunk
constexpr uint32_t WL_STAGE = WL_A_BYTES + WL_B1_BYTES;       // 48 KB
constexpr uint32_t WL_A1_BYTES = WL_KC2 * WL_A_BYTES;         // 128 KB: the activation tile as layer 2's A operand
constexpr uint32_t WL_B2_BYTES = WL_N2 * WL_BK * 2;           // 20 KB
constexpr uint32_t WL_STAGE0 = WL_A1_BYTES / 2;               // stages start behind a1 chunks 0-3
constexpr uint32_t WL_TILE_BYTES = WL_STAGE0 + WL_STAGES * WL_STAGE;  // 208 KB >= a1 tile + W2 ring (168 KB)
static_assert(WL_TILE_BYTES >= WL_A1_BYTES + 2 * WL_B2_BYTES, "a1 chunks 4-7 and the W2 ring overlay the stages");
constexpr uint32_t WL_SMEM = 1024 + WL_TILE_BYTES + (WL_N1 + WL_N2) * 4 + 192;  // tile + biases + 17 mbarriers / the TMEM slot

struct WideArgs {
    const float *b1, *b2;
    __nv_bfloat16 *z2;
    long long m;
    float slope;
    int n_tiles;
};

__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && clock64() - t0 > 20000000000ll) __trap();  // ~10 s: a protocol error must not hang the device
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 x bf16 -> fp32, issued by one thread for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all tcgen05.mma issued so far by this thread -> one arrival on `bar` when they have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor of a K-major bf16 tile stored as rows of 128 bytes (64 elements) with the 128-byte swizzle
// (what TMA's CU_TENSOR_MAP_SWIZZLE_128B writes): start address >> 4, leading byte offset (unused for swizzled K-major) 1,
// stride byte offset 1024 (8 rows), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.  Tile bases are 1024-byte
// aligned (base offset 0); a k-step of 16 elements advances the start address by 32 bytes inside the swizzle atom.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: D fp32 (bits 4-5 = 1), A / B bf16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(WL_THREADS, 1)
wide_layers_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_w1,
                   const __grid_constant__ CUtensorMap map_w2, const WideArgs p) {
    extern __shared__ char wl_raw[];
    char *tile = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(wl_raw) + 1023) & ~(uintptr_t)1023);
    float *bias1 = reinterpret_cast<float *>(tile + WL_TILE_BYTES);
    float *bias2 = bias1 + WL_N1;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias2 + WL_N2);
    uint64_t *full = bars, *empty = bars + 3, *w2_full = bars + 6, *w2_empty = bars + 8;
    uint64_t *d1_full = bars + 10 /* [2]: per N-half */, *a1_ready = bars + 12 /* [2] */, *d2_full = bars + 14, *d2_drained = bars + 15;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < WL_STAGES; i++) mbar_init(full + i, 1), mbar_init(empty + i, 1);
        for (int i = 0; i < 2; i++) mbar_init(w2_full + i, 1), mbar_init(w2_empty + i, 1), mbar_init(d1_full + i, 1), mbar_init(a1_ready + i, 256);
        mbar_init(d2_full, 1), mbar_init(d2_drained, 256);
        prefetch_tmap(&map_a0), prefetch_tmap(&map_w1), prefetch_tmap(&map_w2);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < WL_N1; i += WL_THREADS) bias1[i] = p.b1[i];
    for (int i = threadIdx.x; i < WL_N2; i += WL_THREADS) bias2[i] = p.b2[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            // ---- TMA producer: tiles blockIdx.x, + gridDim.x, ...; the rings' use counters run on across tiles ----
            int it1 = 0, it2 = 0;
            for (int tile_i = blockIdx.x, ti = 0; tile_i < p.n_tiles; tile_i += gridDim.x, ti++) {
                const int m0 = tile_i * WL_BM;
                if (ti) mbar_wait_bounded(d2_full, (ti - 1) & 1);  // the previous tile's layer 2 no longer reads what overlays the stages
                for (int i = 0; i < 2 * WL_KC; i++, it1++) {  // N-half h = i / 8, k-chunk c = i % 8
                    const int h = i / WL_KC, c = i % WL_KC, s = it1 % WL_STAGES, u = it1 / WL_STAGES;
                    mbar_wait_bounded(empty + s, (u & 1) ^ 1);
                    char *st = tile + WL_STAGE0 + s * WL_STAGE;
                    mbar_expect_tx(full + s, WL_STAGE);
                    tma_load_2d(st, &map_a0, c * WL_BK, m0, full + s);
                    tma_load_2d(st + WL_A_BYTES, &map_w1, c * WL_BK, 256 * h, full + s);
                }
                mbar_wait_bounded(d1_full + 1, ti & 1);  // layer 1 has finished reading the stages: W2's ring may overlay them
                for (int c = 0; c < WL_KC2; c++, it2++) {
                    const int s = it2 & 1, u = it2 >> 1;
                    mbar_wait_bounded(w2_empty + s, (u & 1) ^ 1);
                    mbar_expect_tx(w2_full + s, WL_B2_BYTES);
                    tma_load_2d(tile + WL_A1_BYTES + s * WL_B2_BYTES, &map_w2, c * WL_BK, 0, w2_full + s);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issue ----
            constexpr uint32_t ID1 = umma_idesc_bf16(WL_BM, 256), ID2 = umma_idesc_bf16(WL_BM, WL_N2);
            int it1 = 0, it2 = 0;
            for (int tile_i = blockIdx.x, ti = 0; tile_i < p.n_tiles; tile_i += gridDim.x, ti++) {
                if (ti) {  // the previous tile's second epilogue has read its accumulator out of TMEM columns 0..159
                    mbar_wait_bounded(d2_drained, (ti - 1) & 1);
                    tc_fence_after();
                }
                for (int i = 0; i < 2 * WL_KC; i++, it1++) {
                    const int h = i / WL_KC, c = i % WL_KC, s = it1 % WL_STAGES, u = it1 / WL_STAGES;
                    mbar_wait_bounded(full + s, u & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(tile + WL_STAGE0 + s * WL_STAGE);
                    const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(a_addr + WL_A_BYTES);
#pragma unroll
                    for (int k = 0; k < WL_BK / 16; k++) umma_bf16(tmem + 256 * h, da + 2 * k, db + 2 * k, ID1, (c | k) != 0);
                    umma_commit(empty + s);
                    if (c == WL_KC - 1) umma_commit(d1_full + h);  // this half of the accumulator is complete
                }
                for (int c = 0; c < WL_KC2; c++, it2++) {
                    if (c == 0 || c == WL_KC2 / 2) {  // a1 chunks 0-3 come from the first half's epilogue, 4-7 from the second's
                        mbar_wait_bounded(a1_ready + (c ? 1 : 0), ti & 1);
                        tc_fence_after();
                    }
                    const int s = it2 & 1, u = it2 >> 1;
                    mbar_wait_bounded(w2_full + s, u & 1);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(smem_u32(tile + c * WL_A_BYTES));
                    const uint64_t db = umma_desc_sw128(smem_u32(tile + WL_A1_BYTES + s * WL_B2_BYTES));
#pragma unroll
                    for (int k = 0; k < WL_BK / 16; k++) umma_bf16(tmem, da + 2 * k, db + 2 * k, ID2, (c | k) != 0);
                    umma_commit(w2_empty + s);
                }
                umma_commit(d2_full);
            }
        }
    } else {
        // ---- epilogue warps: TMEM lane quadrant q, this thread's row r of the tile; the two warps of a quadrant split the columns ----
        const int q = warp & 3, g = (warp - 2) >> 2, r = q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
        for (int tile_i = blockIdx.x, ti = 0; tile_i < p.n_tiles; tile_i += gridDim.x, ti++) {
        const int m0 = tile_i * WL_BM;
        for (int h = 0; h < 2; h++) {
            mbar_wait_bounded(d1_full + h, ti & 1);
            tc_fence_after();
            for (int jj = 0; jj < 4; jj++) {  // 32 columns at a time: bias, leaky ReLU, bf16, 64 bytes into the swizzled a1 tile
                const int j = 8 * h + 4 * g + jj;
                uint32_t v[32];
                tmem_ld32(lane_addr + 32 * j, v);
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float x0 = __uint_as_float(v[2 * i]) + bias1[32 * j + 2 * i], x1 = __uint_as_float(v[2 * i + 1]) + bias1[32 * j + 2 * i + 1];
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(fmaxf(x0, x0 * p.slope), fmaxf(x1, x1 * p.slope));
                    w[i] = *reinterpret_cast<const uint32_t *>(&hh);
                }
                char *rowp = tile + (j >> 1) * WL_A_BYTES + r * 128;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int unit = ((j & 1) * 4 + i) ^ (r & 7);  // 128-byte swizzle: 16-byte unit index XOR (row mod 8)
                    *reinterpret_cast<uint4 *>(rowp + unit * 16) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
                }
            }
            tc_fence_before();
            fence_async_smem();  // generic-proxy stores above -> visible to the tensor core's (async proxy) reads
            mbar_arrive(a1_ready + h);
        }
        mbar_wait_bounded(d2_full, ti & 1);
        tc_fence_after();
        const long long row = (long long)m0 + r;
        for (int j = g ? 3 : 0; j < (g ? WL_N2 / 32 : 3); j++) {
            uint32_t v[32];
            tmem_ld32(lane_addr + 32 * j, v);
            if (row < p.m) {
                uint4 *dst = reinterpret_cast<uint4 *>(p.z2 + row * WL_N2 + 32 * j);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int c0 = 8 * i + 2 * e;
                        const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(v[c0]) + bias2[32 * j + c0],
                                                                        __uint_as_float(v[c0 + 1]) + bias2[32 * j + c0 + 1]);
                        w[e] = *reinterpret_cast<const uint32_t *>(&hh);
                    }
                    dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tc_fence_before();
        mbar_arrive(d2_drained);  // TMEM columns 0..159 may be overwritten by the next tile's first layer
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

static int make_tmap_bf16(CUtensorMap *m, const void *base, long long rows, int cols, int box_rows) {
    encode_tiled_fn enc = get_encode_fn();
    if (!enc) return fail(OVC_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)WL_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(OVC_E_CUDA, "cuTensorMapEncodeTiled failed", (long long)r);
    return OVC_OK;
}

static int wide_layers_impl(const void *a0, long long m, int k0, const void *w1, const float *b1, int n1, const void *w2, const float *b2,
                            int n2, float slope, void *z2, cudaStream_t st) {
    if (!a0 || !w1 || !b1 || !w2 || !b2 || !z2) return fail(OVC_E_BADARG, "null pointer argument");
    if (k0 != WL_K0 || n1 != WL_N1 || n2 != WL_N2) return fail(OVC_E_UNSUPPORTED, "wide_layers: built for 512 -> 512 -> 160", k0 * 1000000ll + n1 * 1000 + n2);
    if ((((uintptr_t)a0 | (uintptr_t)w1 | (uintptr_t)w2 | (uintptr_t)z2) & 15) != 0) return fail(OVC_E_BADARG, "operands must be 16-byte aligned");
    if (!(slope >= 0.f && slope <= 1.f)) return fail(OVC_E_BADARG, "negative slope must lie in [0, 1]");
    if (m < 0) return fail(OVC_E_BADARG, "negative row count");
    if (m == 0) return OVC_OK;
    CUtensorMap ma, mw1, mw2;
    int rc = make_tmap_bf16(&ma, a0, m, WL_K0, WL_BM);
    if (!rc) rc = make_tmap_bf16(&mw1, w1, WL_N1, WL_K0, 256);  // one N-half of a k-chunk per box
    if (!rc) rc = make_tmap_bf16(&mw2, w2, WL_N2, WL_N1, WL_N2);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(wide_layers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WL_SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "wide_layers kernel attribute");
    WideArgs p;
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    p.b1 = b1, p.b2 = b2, p.z2 = (__nv_bfloat16 *)z2, p.m = m, p.slope = slope;
    p.n_tiles = (int)((m + WL_BM - 1) / WL_BM);
    // persistent: one CTA per SM walks tiles blockIdx.x, + gridDim.x, ... (barriers, TMEM and biases are set up once)
    wide_layers_kernel<<<(unsigned)(p.n_tiles < n_sm ? p.n_tiles : n_sm), WL_THREADS, WL_SMEM, st>>>(ma, mw1, mw2, p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "wide_layers kernel launch");
    return OVC_OK;
}

}  // namespace ovc
