// tcgen05/TMEM implicit-GEMM kernel for the UNet / TAESD convolutions (3x3 stride-1 pad-1, 1x1) and all
// nn.Linear projections of the hot path.
//
// One CTA computes a 128 x block_n output tile. The 128 rows are a BW x BH x BN rectangle of NHWC pixels, so
// every 3x3 tap is ONE 4-D TMA box load at a shifted (w, h) coordinate; TMA's out-of-bounds zero fill is the
// convolution padding. Operands land in shared memory in the 128-byte swizzled K-major layout that
// tcgen05.mma reads directly; accumulators live in TMEM and are read back with tcgen05.ld by 4 epilogue warps
// that fuse bias, time-embedding broadcast, residual add, ReLU / GEGLU / quick-GELU, stage the tile in swizzled shared
// memory and store it with cp.async.bulk.tensor (direct st.global for fp32 / unaligned outputs).
//
// Variants chosen per shape by the engine's autotuner: halo tiles for 3x3 (three column-shifted 8x18 halo tiles per channel
// block serve all nine taps), CTA pairs (tcgen05.mma.cta_group::2 on 256-row tiles, each SM streams half of the weight tile),
// 1-4 k-blocks per pipeline stage, 1-2 CTAs per SM, split-K (+ splitk_reduce_kernel), and conv_persist_kernel (weights
// stationary in shared memory, one CTA per SM) for the TAESD-shaped layers.
//
// warp 0: TMA producer | warp 1: TMEM alloc + MMA issue | warps 2..5: epilogue (TMEM lane quarter = warp % 4)
#include "tc_common.cuh"
#include "vsd_internal.h"
#include <cudaTypedefs.h>
#include <mutex>
#include <vector>

namespace vsd {

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;
static constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB per stage
static constexpr int kGemmThreads = 192;
static constexpr int kHaloABytes = 18 * 1024;          // halo mode: 18 image rows x 8 pixels x 128 B

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Applies bias / per-image vector / residual / ReLU to 32 consecutive accumulator columns and stores them.
// Every loop is fully unrolled with compile-time indices so v[] stays in registers.
//   sbias : optional shared-memory copy of (bias [+ rowvec]) for this tile's columns (index 0 = this chunk)
//   rpre  : optional residual values for this chunk, already loaded (4 x uint4 = 32 bf16)
template <bool kMisc = true>   // kMisc: out_scale / fp32 residual / quick-GELU compiled in (CLIP, AutoencoderKL); the UNet / TAESD GEMMs never use them
__device__ __forceinline__ void epilogue_math32(const GemmParams& p, float (&v)[32], int n_img, long grow, int col,
                                                int ncols, const float* sbias, bool rowvec_in_sbias,
                                                const uint4* rpre) {
    if (ncols <= 0) return;
    const bool full = (ncols >= 32);
    if (sbias) {
        const float4* b4 = reinterpret_cast<const float4*>(sbias);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 t = b4[q];
            v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
        }
    } else if (p.bias) {
        if (full && ((col & 3) == 0)) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 t = __ldg(b4 + q);
                v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += __ldg(p.bias + col + j);
        }
    }
    if (p.rowvec && !rowvec_in_sbias) {
        const float* rv = p.rowvec + (long)n_img * p.N + col;
        if (full && ((col & 3) == 0) && ((p.N & 3) == 0)) {
            const float4* r4 = reinterpret_cast<const float4*>(rv);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 t = __ldg(r4 + q);
                v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += __ldg(rv + j);
        }
    }
    if (kMisc && p.out_scale) {
        const float sc = __ldg(p.out_scale);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sc;
    }
    if (kMisc && p.residual && p.res_f32) {
        const float* r = reinterpret_cast<const float*>(p.residual) + grow * p.ldr + col;
        if (full && ((p.ldr & 3) == 0) && ((col & 3) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(r + q * 4);
                v[q * 4] += t.x; v[q * 4 + 1] += t.y; v[q * 4 + 2] += t.z; v[q * 4 + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += r[j];
        }
    } else if (p.residual) {
        const bf16* r = p.residual + grow * p.ldr + col;
        if (rpre || (full && ((p.ldr & 7) == 0) && ((col & 7) == 0))) {
            const uint4* r4 = reinterpret_cast<const uint4*>(r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 t = rpre ? rpre[q] : r4[q];
                const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
                v[q * 8 + 0] += a.x; v[q * 8 + 1] += a.y; v[q * 8 + 2] += b.x; v[q * 8 + 3] += b.y;
                v[q * 8 + 4] += c.x; v[q * 8 + 5] += c.y; v[q * 8 + 6] += d.x; v[q * 8 + 7] += d.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += __bfloat162float(r[j]);
        }
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (kMisc && p.act == ACT_QUICK_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] / (1.f + __expf(-1.702f * v[j]));
    }
}

template <bool kMisc = true>
__device__ __forceinline__ void epilogue_store32(const GemmParams& p, float (&v)[32], int n_img, long grow, int col,
                                                 int ncols, const float* sbias, bool rowvec_in_sbias,
                                                 const uint4* rpre) {
    if (ncols <= 0) return;
    const bool full = (ncols >= 32);
    epilogue_math32<kMisc>(p, v, n_img, grow, col, ncols, sbias, rowvec_in_sbias, rpre);
    if (p.out_f32) {
        float* o = reinterpret_cast<float*>(p.out) + grow * p.ldo + col;
        if (full && ((p.ldo & 3) == 0) && ((col & 3) == 0)) {
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int q = 0; q < 8; ++q) o4[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) o[j] = v[j];
        }
    } else {
        bf16* o = reinterpret_cast<bf16*>(p.out) + grow * p.ldo + col;
        if (full && ((p.ldo & 7) == 0) && ((col & 7) == 0)) {
            uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                o4[q] = make_uint4(pack_bf16x2(v[q * 8], v[q * 8 + 1]), pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]),
                                   pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]), pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) o[j] = __float2bfloat16(v[j]);
        }
    }
}

// Epilogue of 4 consecutive columns of one output row (bias, per-image row, scale, residual, ReLU, store).
__device__ __forceinline__ void splitk_finish4(const GemmParams& p, float (&v)[4], long grow, int col) {
    const int ncols = min(4, p.N - col);
    const bool vec = (ncols == 4) && ((p.N & 3) == 0);
    if (vec && !p.rowvec && !p.out_scale && !p.res_f32 && (!p.residual || (p.ldr & 3) == 0)) {
        if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        if (p.residual) {
            const uint2 r = *reinterpret_cast<const uint2*>(p.residual + grow * p.ldr + col);
            const float2 a = unpack_bf16x2(r.x), c = unpack_bf16x2(r.y);
            v[0] += a.x; v[1] += a.y; v[2] += c.x; v[3] += c.y;
        }
        if (p.relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
        }
    } else
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < ncols) {
            if (p.bias) v[j] += __ldg(p.bias + col + j);
            if (p.rowvec) v[j] += __ldg(p.rowvec + (grow / ((long)p.H * p.W)) * p.N + col + j);
            if (p.out_scale) v[j] *= __ldg(p.out_scale);
            if (p.residual) v[j] += p.res_f32 ? reinterpret_cast<const float*>(p.residual)[grow * p.ldr + col + j]
                                              : __bfloat162float(p.residual[grow * p.ldr + col + j]);
            if (p.relu) v[j] = fmaxf(v[j], 0.f);
        }
    }
    if (p.out_f32) {
        float* o = reinterpret_cast<float*>(p.out) + grow * p.ldo + col;
        if (vec && ((p.ldo & 3) == 0)) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < ncols) o[j] = v[j];
        }
    } else {
        bf16* o = reinterpret_cast<bf16*>(p.out) + grow * p.ldo + col;
        if (vec && ((p.ldo & 3) == 0)) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
        else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < ncols) o[j] = __float2bfloat16(v[j]);
        }
    }
}

// kPair: the kernel runs as CTA pairs (cluster (2,1,1) over adjacent 128-row tiles): ONE tcgen05.mma.cta_group::2 of the
// leader drives both SMs' tensor cores on a 256 x block_n tile, each CTA streams its own 128 activation rows but only
// HALF of the weight tile (block_n / 2 rows) -- the weight bytes entering each SM are halved.
// kFeat: epilogue features compiled into this instantiation (bit 0 GEGLU, bit 1 folded-LayerNorm consumer, bit 2 row statistics
// for a LayerNorm consumer, bit 3 out_scale / fp32 residual / quick-GELU). The full kernel is ~8 300 SASS instructions (133 KB,
// more than the SM's instruction cache) and ncu shows the epilogue warps of these 10-20 us kernels stalled on instruction fetch
// (`no_inst` = 27 % of their samples, profiles/r02_summary.md); most layers need none of the features, so they run a lean
// instantiation. launch_gemm_op picks the smallest instantiated superset.
enum { FEAT_GEGLU = 1, FEAT_LN = 2, FEAT_STATS = 4, FEAT_MISC = 8, FEAT_ALL = 15 };
template <int kEpi, bool kPair, int kFeat>   // kEpi 0: direct st.global epilogue, 1: bf16 tile through smem + TMA store, 2: fp32 split-K partials through TMA,
                      // 3: split-K across a thread-block cluster, reduced through distributed shared memory
__global__ void __launch_bounds__(kGemmThreads, ((kEpi == 1 || kEpi == 3) && !kPair ? 2 : 1))
conv_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapR, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    const int kbs = p.kb_per_stage;                       // 64-wide k-blocks carried by one pipeline stage
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs, owns the full barriers)
    const uint32_t b_rows = kPair ? (uint32_t)p.block_n >> 1 : (uint32_t)p.block_n;   // weight rows this CTA streams
    const uint32_t b_bytes = b_rows * 128u;               // one k-block of (this CTA's part of) the B tile
    // stage layout: [A k-block 0 .. kbs-1][B k-block 0 .. kbs-1]; halo mode: [A halo tile 8 x 18 px][B tap dy=-1,0,1]
    const bool halo = p.halo != 0;
    const uint32_t a_stage = halo ? (uint32_t)kHaloABytes : (uint32_t)kbs * kABytes;
    const uint32_t stage_bytes = halo ? ((uint32_t)kHaloABytes + 3u * b_bytes) : (uint32_t)kbs * ((uint32_t)kABytes + b_bytes);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* empty_bar = full_bar + stages;
    uint64_t* tmem_full_bar = empty_bar + stages;
    uint64_t* res_bar = tmem_full_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);
    uint8_t* stag = smem + p.stage_off;   // epilogue staging (and residual tile) region
    // [block_n] bias (+ time-embedding row) of this tile, 16-byte aligned for float4 reads
    float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
    float* swsum = sbias + p.block_n;          // LayerNorm fold: [block_n] column sums of W' (mode 1) | [block_n] column means (mode 2)
    float* scol = swsum + p.block_n;           // mode 2: [block_n] column rstd

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr bool fGeglu = (kFeat & FEAT_GEGLU) != 0, fMisc = (kFeat & FEAT_MISC) != 0;
    const bool geglu = fGeglu && p.act == ACT_GEGLU;
    const int ln_mode = (kFeat & FEAT_LN) ? p.ln_mode : 0;
    float* const rowstats_out = (kFeat & FEAT_STATS) ? p.rowstats_out : nullptr;
    // phase stamps (tools/gemm_phase_timing.py) are compiled in only with -DVSD_GEMM_STAMPS: they sit in the hot loops
#ifdef VSD_GEMM_STAMPS
    const bool dbg_cta = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#else
    constexpr bool dbg_cta = false;
#endif
#define VSD_STAMP(i) do { if (dbg_cta) { p.dbg[i] = clock64(); p.dbg[100 + (i)] = (long long)globaltimer_ns(); } } while (0)
    if (threadIdx.x == 0) VSD_STAMP(0);
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace, 1u);

    const int mt = blockIdx.x;
    const int tw = mt % p.tiles_w;
    const int th = (mt / p.tiles_w) % p.tiles_h;
    const int tn = mt / (p.tiles_w * p.tiles_h);
    const int w0 = tw * p.BW, h0 = th * p.BH, n0 = tn * p.BN;
    const int col0 = blockIdx.y * p.block_n;
    const int bcol0 = col0 + (int)(rank * b_rows);       // first weight row this CTA loads
    const int split = blockIdx.z;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(kb_begin + p.kb_per_split, p.kb_total);

    // Stage bookkeeping shared by the early prefetch and the producer loop
    const int total_iters = halo ? (kb_end - kb_begin) : (kb_end - kb_begin + kbs - 1) / kbs;
    const int npre = ((p.a_static && !halo) || p.b_static) ? min(stages, total_iters) : 0;
    // Loads count their bytes on the LEADER's full barrier (pair mode: through its shared::cluster address); only the leader
    // arms it, with the bytes of both CTAs.
    const uint32_t fbar_addr = kPair ? dsmem_addr(smem_u32(full_bar), 0u) : smem_u32(full_bar);
    auto load_a = [&](void* dst, int s, int c0, int c1, int c2, int c3) {
        if (kPair) tma_load_4d_pair(dst, &mapA, fbar_addr + (uint32_t)s * 8u, c0, c1, c2, c3);
        else tma_load_4d(dst, &mapA, &full_bar[s], c0, c1, c2, c3);
    };
    auto load_b = [&](void* dst, int s, int c0, int c1) {
        if (kPair) tma_load_2d_pair(dst, &mapB, fbar_addr + (uint32_t)s * 8u, c0, c1);
        else tma_load_2d(dst, &mapB, &full_bar[s], c0, c1);
    };
    auto expect = [&](int s, uint32_t bytes) {
        if (!kPair) mbar_expect_tx(&full_bar[s], bytes);
        else if (rank == 0) mbar_expect_tx(&full_bar[s], 2u * bytes);
    };
    // The constant operand (weights) of the first ring pass is requested right away: before waiting for the producer
    // kernel of the activations (PDL) and, without pairs, even before the TMEM allocation / CTA barrier.
    auto early_prefetch = [&]() {
        int kb = kb_begin;
        if (halo) {
            const int kpt = p.cin >> 6;
            for (int it = 0; it < npre; ++it, ++kb) {   // iteration = (channel block, column shift)
                uint8_t* sb = smem + (size_t)it * stage_bytes + a_stage;
                const int cb = kb / 3, dxi = kb - cb * 3;
                expect(it, stage_bytes);
                for (int dyi = 0; dyi < 3; ++dyi)
                    load_b(sb + (size_t)dyi * b_bytes, it, ((dyi * 3 + dxi) * kpt + cb) * 64, bcol0);
            }
        } else {
            for (int it = 0; it < npre; ++it) {
                const int nkb = min(kbs, kb_end - kb);
                uint8_t* sa = smem + (size_t)it * stage_bytes;
                uint8_t* sb = sa + a_stage;
                expect(it, (uint32_t)nkb * ((uint32_t)kABytes + b_bytes));
                for (int j = 0; j < nkb; ++j, ++kb) {
                    if (p.b_static) load_b(sb + (size_t)j * b_bytes, it, kb * 64, bcol0);
                    else load_a(sa + (size_t)j * kABytes, it, kb * 64, w0, h0, n0);  // taps == 1
                }
            }
        }
    };
    if (warp == 0) {
        if (elect_one()) {
            tma_prefetch_desc(&mapA);
            tma_prefetch_desc(&mapB);
            for (int s = 0; s < stages; ++s) {
                mbar_init(&full_bar[s], 1);
                mbar_init(&empty_bar[s], 1);
            }
            mbar_init(tmem_full_bar, 1);
            mbar_init(res_bar, 1);
            fence_barrier_init();
            if (kEpi != 0) tma_prefetch_desc(&mapC);
            if (kEpi == 1 && p.tma_res) tma_prefetch_desc(&mapR);
            if (!kPair) early_prefetch();   // pairs: after the cluster barrier (the peer's barriers must exist first)
        }
        __syncwarp();
    }
    if (kPair) {
        // Both CTAs of the pair must be running before tcgen05.alloc.cta_group::2 reaches into the peer SM's tensor memory.
        // The two CTAs of a cluster are co-scheduled but do not start at the same instant; with several lanes and PDL keeping
        // every SM busy the skew grows, and an allocation issued before the peer CTA had started never returned (VSD_TRACE:
        // pairs that entered the kernel and never owned their columns while nothing else was running on the GPU).
        cluster_sync_all();
        if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 3, 1u);
        if (warp == 0) {
            if (elect_one()) early_prefetch();            // the peer's barriers exist now
            __syncwarp();
        }
        if (warp == 1) tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    }
    tc_fence_before_sync();
    if (kPair) cluster_sync_all(); else __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // The next kernel may begin its own prologue / weight prefetch now -- but only now that this CTA OWNS its TMEM columns.
    // Triggered before the allocation (as in round 1), CTAs of the dependent kernel could become co-resident and take columns
    // first: capacity is guaranteed (shared-memory guard in build_gemm_op), contiguity is not, and a dependent that fragments
    // the free space (or the common range a cta_group::2 allocation needs on both SMs) waits for this kernel, which waits for
    // its columns: a rare, timing-dependent deadlock (seen once in a single-engine run and once with two batch-4 lanes; the
    // device watchdog caught the second). With the trigger after the allocation every TMEM wait is on an older kernel.
    pdl_launch_dependents();
    if (threadIdx.x == 0) VSD_STAMP(1);
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 1, 1u);

    if (warp == 0) {
        // One elected thread runs the whole producer loop (single active thread => ptxas keeps the loop state in
        // uniform registers and issues UTMALDG directly).
        if (elect_one()) {
            const int kpt = p.cin >> 6;  // 64-channel blocks per tap
            int tap = kb_begin / kpt;
            int cb = kb_begin - tap * kpt;
            int s = 0, dbg_it = 0;
            uint32_t ph = 0;
            pdl_wait();
            VSD_STAMP(7);
            // The residual tile rides behind the first ring pass: [chunk of 32 columns][128 rows][64 B], 64-byte swizzle
            const int res_at = min(stages, total_iters) - 1;
            auto issue_residual = [&]() {
                mbar_expect_tx(res_bar, (uint32_t)p.block_n * 256u);
                for (int c = 0; c < p.block_n; c += 32)
                    tma_load_4d(stag + (size_t)(c >> 5) * 8192, &mapR, res_bar, col0 + c, w0, h0, n0);
            };
            if (halo) {
                // iteration = (64-channel block cb, column shift dx): ONE 8 x 18-pixel halo tile serves the three
                // row taps (dy = -1, 0, +1) as 1024-byte-aligned offsets; three weight tiles (one per dy) ride along.
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    const bool pre = dbg_it < npre;
                    if (!pre) {
                        mbar_wait(&empty_bar[s], ph ^ 1u, 1);
                        expect(s, stage_bytes);
                    }
                    uint8_t* sa = smem + (size_t)s * stage_bytes;
                    uint8_t* sb = sa + a_stage;
                    const int cbh = kb / 3, dxi = kb - cbh * 3;
                    load_a(sa, s, cbh * 64, w0 + dxi - 1, h0 - 1, n0);
                    if (!pre)
                        for (int dyi = 0; dyi < 3; ++dyi)
                            load_b(sb + (size_t)dyi * b_bytes, s, ((dyi * 3 + dxi) * kpt + cbh) * 64, bcol0);
                    if (kEpi == 1 && p.tma_res && dbg_it == res_at) issue_residual();
                    ++dbg_it;
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            } else
            for (int kb = kb_begin; kb < kb_end;) {
                const int nkb = min(kbs, kb_end - kb);
                const bool pre = dbg_it < npre;
                if (!pre) {
                    mbar_wait(&empty_bar[s], ph ^ 1u, 1);
                    expect(s, (uint32_t)nkb * ((uint32_t)kABytes + b_bytes));
                }
                if (dbg_cta && dbg_it < 16) p.dbg[16 + 2 * dbg_it] = clock64();
                uint8_t* sa = smem + (size_t)s * stage_bytes;
                uint8_t* sb = sa + a_stage;
                for (int j = 0; j < nkb; ++j, ++kb) {
                    int dy = 0, dx = 0;
                    if (p.taps == 9) {
                        const int ty = (tap * 11) >> 5;  // tap / 3 for tap in [0, 9)
                        dy = ty - 1;
                        dx = tap - ty * 3 - 1;
                    }
                    if (!(pre && p.a_static))
                        load_a(sa + (size_t)j * kABytes, s, cb * 64, p.cstride * w0 + dx + p.cshift, p.cstride * h0 + dy + p.cshift, n0);
                    if (!(pre && p.b_static))
                        load_b(sb + (size_t)j * b_bytes, s, kb * 64, bcol0);
                    if (++cb == kpt) { cb = 0; ++tap; }
                }
                if (dbg_cta && dbg_it < 16) p.dbg[16 + 2 * dbg_it + 1] = clock64();
                if (kEpi == 1 && p.tma_res && dbg_it == res_at) issue_residual();
                ++dbg_it;
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 1 && (!kPair || rank == 0)) {
        const uint32_t idesc = umma_idesc_bf16(kPair ? 2 * kBlockM : kBlockM, (uint32_t)p.block_n);
        int s = 0, dbg_it = 0;
        uint32_t ph = 0;
        bool first = true;
        for (int kb = kb_begin; kb < kb_end;) {
            const int nkb = halo ? 3 : min(kbs, kb_end - kb);   // halo: three row taps per stage
            mbar_wait(&full_bar[s], ph, 2);
            tc_fence_after_sync();
            if (lane == 0 && first) VSD_STAMP(2);
            if (lane == 0 && dbg_cta && dbg_it < 16) p.dbg[64 + 2 * dbg_it] = clock64();
            {
                const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t b_addr = a_addr + a_stage;
                const uint32_t a_step = halo ? 1024u : (uint32_t)kABytes;   // halo: next row tap = next 8-pixel image row
                if (elect_one()) {   // one elected lane issues; ptxas keeps descriptors in uniform registers
                    for (int j = 0; j < nkb; ++j) {
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            if (kPair)
                                umma_bf16_pair(tmem_base, umma_desc_sw128(a_addr + j * a_step + k * 32),
                                               umma_desc_sw128(b_addr + j * b_bytes + k * 32), idesc,
                                               (first && j == 0 && k == 0) ? 0u : 1u);
                            else
                                umma_bf16(tmem_base, umma_desc_sw128(a_addr + j * a_step + k * 32),
                                          umma_desc_sw128(b_addr + j * b_bytes + k * 32), idesc,
                                          (first && j == 0 && k == 0) ? 0u : 1u);
                        }
                    }
                    if (kPair) umma_commit_pair(&empty_bar[s]);   // frees the stage in both CTAs when these MMAs retire
                    else umma_commit(&empty_bar[s]);
                }
                __syncwarp();
                if (lane == 0 && dbg_cta && dbg_it < 16) p.dbg[64 + 2 * dbg_it + 1] = clock64();
            }
            ++dbg_it;
            first = false;
            kb += halo ? 1 : nkb;
            if (++s == stages) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) {
            if (kPair) umma_commit_pair(tmem_full_bar);
            else umma_commit(tmem_full_bar);
        }
        __syncwarp();
        if (lane == 0) VSD_STAMP(3);
    } else if (warp >= 2) {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int bw = r % p.BW;
        const int bh = (r / p.BW) % p.BH;
        const int bn = r / (p.BW * p.BH);
        const int n_img = n0 + bn, hh = h0 + bh, ww = w0 + bw;
        const bool row_ok = (n_img < p.NB) && (hh < p.H) && (ww < p.W);
        const long grow = ((long)n_img * p.H + hh) * p.W + ww;
        const long rows_total = (long)p.NB * p.H * p.W;
        // While the main loop runs: stage bias (+ the per-image time-embedding row when the tile lies inside one
        // image) in shared memory and prefetch the first residual chunk.
        const bool rowvec_in_sbias = (p.rowvec != nullptr) && (p.BN == 1);
        const bool use_sbias = (p.splits == 1) && (p.bias != nullptr || rowvec_in_sbias);
        if (use_sbias) {
            for (int i = threadIdx.x - 64; i < p.block_n; i += 128) {
                const int col = col0 + i;
                float bv = 0.f;
                if (col < p.N) {
                    if (p.bias) bv = __ldg(p.bias + col);
                    if (rowvec_in_sbias) bv += __ldg(p.rowvec + (long)n0 * p.N + col);
                }
                sbias[i] = bv;
            }
        }
        if (ln_mode == 1)
            for (int i = threadIdx.x - 64; i < p.block_n; i += 128) swsum[i] = (col0 + i < p.N) ? __ldg(p.ln_wsum + col0 + i) : 0.f;
        pdl_wait();   // bias / time-embedding rows above are constants; everything below depends on earlier kernels
        // LayerNorm folded into this GEMM (GemmParams::ln_mode): the GEMM that PRODUCED the normalised operand left, per row
        // and per N tile of its own grid, the partial sums of x and x^2 of the values it stored (rowstats_out below). Reduce
        // them here (fixed order) to mean / rstd: per output row (mode 1) or per output column (mode 2, swapped operands).
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (ln_mode) {
            const float inv_k = 1.0f / (float)(p.cin * p.taps);
            if (ln_mode == 1) {
                if (row_ok) {
                    const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + grow * p.ln_nst;
                    float a = 0.f, b = 0.f;
                    for (int t = 0; t < p.ln_nst; ++t) { const float2 v = st[t]; a += v.x; b += v.y; }
                    ln_mean = a * inv_k;
                    ln_rstd = rsqrtf(fmaxf(b * inv_k - ln_mean * ln_mean, 0.f) + p.ln_eps);
                }
            } else {
                for (int i = threadIdx.x - 64; i < p.block_n; i += 128) {
                    float a = 0.f, b = 0.f;
                    if (col0 + i < p.N) {
                        const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + (long)(col0 + i) * p.ln_nst;
                        for (int t = 0; t < p.ln_nst; ++t) { const float2 v = st[t]; a += v.x; b += v.y; }
                    }
                    const float m = a * inv_k;
                    swsum[i] = m;
                    scol[i] = rsqrtf(fmaxf(b * inv_k - m * m, 0.f) + p.ln_eps);
                }
            }
        }
        const bool res_vec = (p.residual != nullptr) && row_ok && ((p.ldr & 7) == 0) && ((col0 & 7) == 0) &&
                             (p.splits == 1) && !geglu && !p.tma_res && !(fMisc && p.res_f32);
        // this warp's 32 rows as a sub-box of the tile rectangle (all extents are powers of two)
        const int sw0 = w0 + (q * 32) % p.BW, sh0 = h0 + ((q * 32) / p.BW) % p.BH, sn0 = n0 + (q * 32) / (p.BW * p.BH);
        uint4 rnext[4];
        if (res_vec && col0 + 32 <= p.N) {
            const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + grow * p.ldr + col0);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) rnext[q4] = r4[q4];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // epilogue warps only: sbias is complete
        mbar_wait(tmem_full_bar, 0, 3);
        tc_fence_after_sync();
        if (threadIdx.x == 64) VSD_STAMP(4);
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
        if (kEpi == 3) {
            // split-K inside a cluster: nothing to do here, the accumulator is sent to its reducer after the cluster barrier
        } else if (kEpi == 2) {
            // split-K partials: [32 rows][128 B] fp32 chunks, 128-byte swizzle, stored with one 5-D box per warp and chunk
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t u[32];
                tmem_ld32(tbase + c, u);
                uint8_t* buf = stag + (size_t)((c >> 5) * 4 + q) * 4096;
                uint8_t* myrow = buf + lane * 128;
                const int sx = lane & 7;
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(myrow + ((j ^ sx) << 4)) = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (col0 + c < p.N && elect_one()) {
                    tma_store_5d(&mapC, buf, col0 + c, sw0, sh0, sn0, split);
                    tma_store_commit();
                }
            }
            tma_store_wait_all();
        } else if (kEpi == 1 && geglu) {
            const int half = p.block_n >> 1;
            const int ocol0 = col0 >> 1;
            for (int c = 0; c < half; c += 32) {
                // the weight rows are interleaved in groups of 128 = 64 value rows + their 64 gate rows (at load time), so a
                // 128- or 256-column tile holds whole groups: output column c pairs tile columns vc and vc + 64
                const int vc = ((c >> 6) << 7) + (c & 63);
                uint32_t u[32], g[32];
                tmem_ld32(tbase + vc, u);
                tmem_ld32(tbase + vc + 64, g);
                uint8_t* buf = stag + (size_t)((c >> 5) * 4 + q) * 2048;
                uint8_t* myrow = buf + lane * 64;
                const int sx = (lane >> 1) & 3;
                tmem_ld_wait();
                const float4* bu = reinterpret_cast<const float4*>(sbias + vc);
                const float4* bg = reinterpret_cast<const float4*>(sbias + vc + 64);
                if (ln_mode == 1) {          // folded LayerNorm (norm3 -> GEGLU projection)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        u[j] = __float_as_uint(ln_rstd * (__uint_as_float(u[j]) - ln_mean * swsum[vc + j]));
                        g[j] = __float_as_uint(ln_rstd * (__uint_as_float(g[j]) - ln_mean * swsum[vc + 64 + j]));
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v[8];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const float4 tu = bu[j * 2 + h2], tg = bg[j * 2 + h2];
                        const int b0 = j * 8 + h2 * 4;
                        v[h2 * 4 + 0] = (__uint_as_float(u[b0 + 0]) + tu.x) * gelu_erf(__uint_as_float(g[b0 + 0]) + tg.x);
                        v[h2 * 4 + 1] = (__uint_as_float(u[b0 + 1]) + tu.y) * gelu_erf(__uint_as_float(g[b0 + 1]) + tg.y);
                        v[h2 * 4 + 2] = (__uint_as_float(u[b0 + 2]) + tu.z) * gelu_erf(__uint_as_float(g[b0 + 2]) + tg.z);
                        v[h2 * 4 + 3] = (__uint_as_float(u[b0 + 3]) + tu.w) * gelu_erf(__uint_as_float(g[b0 + 3]) + tg.w);
                    }
                    *reinterpret_cast<uint4*>(myrow + ((j ^ sx) << 4)) =
                        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one()) {
                    tma_store_4d(&mapC, buf, ocol0 + c, sw0, sh0, sn0);
                    tma_store_commit();
                }
            }
            tma_store_wait_all();
        } else if (kEpi == 1) {
            const float ln_ws_row = (ln_mode == 2 && row_ok) ? __ldg(p.ln_wsum + grow) : 0.f;
            const float ln_rb_row = (ln_mode == 2 && row_ok && p.ln_rowbias) ? __ldg(p.ln_rowbias + grow) : 0.f;
            float rst_s = 0.f, rst_q = 0.f;
            if (p.tma_res) mbar_wait(res_bar, 0, 4);
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t u[32];
                tmem_ld32(tbase + c, u);
                const int col = col0 + c;
                const int ncols = min(32, p.N - col);
                uint8_t* buf = stag + (size_t)((c >> 5) * 4 + q) * 2048;
                uint8_t* myrow = buf + lane * 64;
                const int sx = (lane >> 1) & 3;
                uint4 rcur[4];
                const bool have_pre = res_vec && ncols == 32;
                if (have_pre) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) rcur[q4] = rnext[q4];
                    if (c + 32 < p.block_n && col + 64 <= p.N) {
                        const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + grow * p.ldr + col + 32);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) rnext[q4] = r4[q4];
                    }
                }
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]);
                if (ln_mode == 1) {          // folded LayerNorm over the rows of A: out = rstd * (acc - mean * colsum(W')) + bias'
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = ln_rstd * (v[j] - ln_mean * swsum[c + j]);
                } else if (ln_mode == 2) {   // ... over the rows of B (output columns); per-row constants of the weight operand
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = scol[c + j] * (v[j] - swsum[c + j] * ln_ws_row) + ln_rb_row;
                }
                if (row_ok) {
                    // one instance of the (fully unrolled) epilogue arithmetic: the residual chunk comes from the TMA-staged tile
                    // or from the prefetched registers
                    if (p.tma_res) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) rcur[j] = *reinterpret_cast<const uint4*>(myrow + ((j ^ sx) << 4));
                    }
                    epilogue_math32<fMisc>(p, v, n_img, grow, col, ncols, use_sbias ? sbias + c : nullptr, rowvec_in_sbias,
                                           (p.tma_res || have_pre) ? rcur : nullptr);
                    if (rowstats_out) {      // a LayerNorm consumes this tensor: partial row sums of what is being stored
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) { rst_s += v[j]; rst_q += v[j] * v[j]; }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(myrow + ((j ^ sx) << 4)) =
                        make_uint4(pack_bf16x2(v[j * 8], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                   pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
                fence_proxy_async_smem();
                __syncwarp();
                if (ncols > 0 && elect_one()) {
                    tma_store_4d(&mapC, buf, col, sw0, sh0, sn0);
                    tma_store_commit();
                }
            }
            if (rowstats_out && row_ok)
                reinterpret_cast<float2*>(rowstats_out)[grow * gridDim.y + blockIdx.y] = make_float2(rst_s, rst_q);
            tma_store_wait_all();
        } else if (kEpi != 0) {
        } else if (p.splits > 1) {
            float* dst = p.partial + ((long)split * rows_total + grow) * p.N;
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t u[32];
                tmem_ld32(tbase + c, u);
                tmem_ld_wait();
                const int col = col0 + c;
                const int ncols = min(32, p.N - col);
                if (row_ok && ncols > 0) {
                    if (ncols == 32 && ((p.N & 3) == 0)) {
                        float4* d4 = reinterpret_cast<float4*>(dst + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            d4[j] = make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]),
                                                __uint_as_float(u[4 * j + 2]), __uint_as_float(u[4 * j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) dst[col + j] = __uint_as_float(u[j]);
                    }
                }
            }
        } else if (geglu) {
            const int half = p.block_n >> 1;
            const int ocol0 = col0 >> 1;
            for (int c = 0; c < half; c += 32) {
                const int vc = ((c >> 6) << 7) + (c & 63);   // see the tensor-store GEGLU epilogue
                uint32_t u[32], g[32];
                tmem_ld32(tbase + vc, u);
                tmem_ld32(tbase + vc + 64, g);
                tmem_ld_wait();
                float v[32];
                const float4* bu = reinterpret_cast<const float4*>(sbias + vc);
                const float4* bg = reinterpret_cast<const float4*>(sbias + vc + 64);
#pragma unroll
                for (int q4 = 0; q4 < 8; ++q4) {
                    const float4 tu = bu[q4], tg = bg[q4];
                    v[q4 * 4 + 0] = (__uint_as_float(u[q4 * 4 + 0]) + tu.x) * gelu_erf(__uint_as_float(g[q4 * 4 + 0]) + tg.x);
                    v[q4 * 4 + 1] = (__uint_as_float(u[q4 * 4 + 1]) + tu.y) * gelu_erf(__uint_as_float(g[q4 * 4 + 1]) + tg.y);
                    v[q4 * 4 + 2] = (__uint_as_float(u[q4 * 4 + 2]) + tu.z) * gelu_erf(__uint_as_float(g[q4 * 4 + 2]) + tg.z);
                    v[q4 * 4 + 3] = (__uint_as_float(u[q4 * 4 + 3]) + tu.w) * gelu_erf(__uint_as_float(g[q4 * 4 + 3]) + tg.w);
                }
                if (row_ok) {
                    bf16* o = reinterpret_cast<bf16*>(p.out) + grow * p.ldo + ocol0 + c;
                    uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq)
                        o4[qq] = make_uint4(pack_bf16x2(v[qq * 8], v[qq * 8 + 1]), pack_bf16x2(v[qq * 8 + 2], v[qq * 8 + 3]),
                                            pack_bf16x2(v[qq * 8 + 4], v[qq * 8 + 5]), pack_bf16x2(v[qq * 8 + 6], v[qq * 8 + 7]));
                }
            }
        } else {
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t u[32];
                tmem_ld32(tbase + c, u);
                const int col = col0 + c;
                const int ncols = min(32, p.N - col);
                uint4 rcur[4];
                const bool have_pre = res_vec && ncols == 32;
                if (have_pre) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) rcur[q4] = rnext[q4];
                    if (c + 32 < p.block_n && col + 64 <= p.N) {   // prefetch the next chunk's residual
                        const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + grow * p.ldr + col + 32);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) rnext[q4] = r4[q4];
                    }
                }
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]);
                if (row_ok)
                    epilogue_store32<fMisc>(p, v, n_img, grow, col, ncols, use_sbias ? sbias + c : nullptr, rowvec_in_sbias,
                                            have_pre ? rcur : nullptr);
            }
        }
    }
    if (kEpi == 3) {
        // Split-K inside a thread-block cluster (1,1,splits): the CTAs are co-scheduled, so they can exchange their fp32
        // accumulators through DISTRIBUTED SHARED MEMORY instead of a workspace in L2 + a second kernel. CTA z reduces the row
        // band [z * rows_per, (z + 1) * rows_per) of the tile: every CTA pushes that band of its accumulator into CTA z's
        // (idle) operand ring with st.shared::cluster, one cluster barrier later CTA z sums the `splits` copies in split order
        // (=> deterministic), applies the epilogue and stores coalesced rows.
        const int rows_per = (kBlockM + p.splits - 1) / p.splits;
        float* recv = reinterpret_cast<float*>(smem);          // [splits][rows_per][block_n] fp32, aliases the operand ring
        // The ring of EVERY CTA must be idle before anyone writes into it: this CTA's last MMAs have retired (tmem_full_bar;
        // its TMA loads completed before those MMAs could be issued).
        if (warp == 1) { mbar_wait(tmem_full_bar, 0, 8); tc_fence_after_sync(); }
        if (threadIdx.x == 64) VSD_STAMP(8);
        cluster_sync_all();
        if (warp >= 2) {
            const int q = warp & 3;
            const int r = q * 32 + lane;                       // tile row held by this thread (TMEM lane)
            const int dst = r / rows_per, lr = r - dst * rows_per;
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
            const uint32_t remote = dsmem_addr(smem_u32(recv + ((size_t)split * rows_per + lr) * p.block_n), (uint32_t)dst);
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t u[32];
                tmem_ld32(tbase + c, u);
                tmem_ld_wait();
                // rows are block_n * 4 bytes apart (a multiple of 128): XOR the 16-byte chunk index with the row so that the
                // eight lanes of a store phase hit different banks of the receiver's shared memory
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};"
                                 ::"r"(remote + ((uint32_t)(((c >> 2) + j) ^ (lr & 7)) << 4)),
                                 "r"(u[4 * j]), "r"(u[4 * j + 1]), "r"(u[4 * j + 2]), "r"(u[4 * j + 3]) : "memory");
            }
        }
        cluster_sync_all();                                    // release / acquire: every band has arrived
        if (threadIdx.x == 64) VSD_STAMP(9);
        if (warp >= 2) {
            // A warp walks rows of this CTA's band, its lanes walk 4-column groups (conflict-free 16-byte smem reads, coalesced
            // global stores). No integer division: tile extents are powers of two (lbw / lbh).
            const int g4 = p.block_n >> 2;
            const int r_begin = split * rows_per;
            const int r_end = min(kBlockM, r_begin + rows_per);
            const int lg = (g4 <= 8) ? 3 : ((g4 <= 16) ? 4 : 5);          // lanes per row = 1 << lg (>= g4 when g4 <= 32)
            const int lanes_row = 1 << lg;
            const int rows_warp = 32 >> lg;                               // rows one warp covers per step
            const int lane_r = lane >> lg, lane_g = lane & (lanes_row - 1);
            const int wq = warp - 2;
            const size_t src_stride4 = ((size_t)rows_per * p.block_n) >> 2;   // float4 units between two splits' copies
            for (int rb = r_begin + wq * rows_warp; rb < r_end; rb += 4 * rows_warp) {   // warp-uniform trip count (shuffles below)
                const int r = rb + lane_r;
                const int n_img = n0 + (r >> (p.lbw + p.lbh)), hh = h0 + ((r >> p.lbw) & (p.BH - 1)), ww = w0 + (r & (p.BW - 1));
                const bool rok = (r < r_end) && (n_img < p.NB) && (hh < p.H) && (ww < p.W);
                const long grow = ((long)n_img * p.H + hh) * p.W + ww;
                const int lr = r - r_begin;
                float rst_s = 0.f, rst_q = 0.f;
                for (int g0 = lane_g; g0 < g4; g0 += lanes_row) {
                    const int col = col0 + g0 * 4;
                    if (!rok || col >= p.N) continue;
                    const float4* src = reinterpret_cast<const float4*>(recv + (size_t)lr * p.block_n) + (g0 ^ (lr & 7));
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
                    for (int z = 0; z < p.splits; ++z) {
                        const float4 t = src[(size_t)z * src_stride4];
                        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                    }
                    splitk_finish4(p, v, grow, col);     // leaves the stored values in v
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (col + j < p.N) { rst_s += v[j]; rst_q += v[j] * v[j]; }
                }
                if (rowstats_out) {                    // LayerNorm consumer downstream: the lanes sharing a row fold their sums
                    for (int o = lanes_row >> 1; o > 0; o >>= 1) {
                        rst_s += __shfl_xor_sync(0xffffffffu, rst_s, o);
                        rst_q += __shfl_xor_sync(0xffffffffu, rst_q, o);
                    }
                    if (rok && lane_g == 0)
                        reinterpret_cast<float2*>(rowstats_out)[grow * gridDim.y + blockIdx.y] = make_float2(rst_s, rst_q);
                }
            }
        }
        if (threadIdx.x == 64) VSD_STAMP(10);
    }
    if (threadIdx.x == 64) VSD_STAMP(5);
    tc_fence_before_sync();
    if (kPair) cluster_sync_all(); else __syncthreads();   // pairs: neither CTA frees TMEM / leaves while the other still uses the pair
    if (warp == 1) {
        if (kPair) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
        else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
    if (threadIdx.x == 32) VSD_STAMP(6);
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 2, 1u);
#undef VSD_STAMP
}

// ------------------------------------------------------------------------------------------ persistent 3x3 convolution
// For the TAESD layers (64 -> <= 64 channels on up to 512 x 512 pixels) the whole weight tensor is 9 x N x 64 x 2 B <= 72 KiB, yet
// conv_gemm_kernel re-fetches it for every 128-pixel tile (57 % of the bytes entering the SM). Here one CTA per SM keeps the
// weights in shared memory and walks over the 8 x 16-pixel tiles: per tile only the three column-shifted 8 x 18 halo tiles of
// activations are loaded; two TMEM accumulators alternate so the epilogue of tile i overlaps the MMAs of tile i + 1.
// warp 0: TMA producer | warp 1: MMA issue (+ TMEM alloc) | warps 2..5: epilogue (bias / residual / ReLU -> smem -> TMA store)
__global__ void __launch_bounds__(kGemmThreads, 1)
conv_persist_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const __grid_constant__ CUtensorMap mapC, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int S = p.stages;                                   // activation ring depth
    const uint32_t b_bytes = (uint32_t)p.block_n * 128u;      // one tap of the weights: [block_n][64] bf16
    uint8_t* sB = smem;                                       // [9 taps][b_bytes]
    uint8_t* sA = sB + 9u * b_bytes;                          // [S][18 KiB] halo tiles
    uint8_t* stag = sA + (size_t)S * kHaloABytes;             // [block_n / 32][4 warps][2 KiB] output staging
    uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + p.bar_off);
    uint64_t* a_full = b_full + 1;
    uint64_t* a_empty = a_full + S;
    uint64_t* acc_full = a_empty + S;                         // [2]
    uint64_t* acc_empty = acc_full + 2;                       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* sbias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace, 1u);
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        tma_prefetch_desc(&mapC);
        mbar_init(b_full, 1);
        for (int s = 0; s < S; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int k = 0; k < 2; ++k) { mbar_init(&acc_full[k], 1); mbar_init(&acc_empty[k], 128); }
        fence_barrier_init();
        mbar_expect_tx(b_full, 9u * b_bytes);                 // the weights are constants: fetch them before the PDL wait
        for (int tap = 0; tap < 9; ++tap) tma_load_2d(sB + (size_t)tap * b_bytes, &mapB, b_full, tap * 64, 0);
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    for (int i = threadIdx.x; i < p.block_n; i += blockDim.x) sbias[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.f;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();                                  // after the TMEM allocation: see conv_gemm_kernel
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 1, 1u);
    const uint32_t acc_cols = (uint32_t)p.tmem_cols >> 1;     // columns of one accumulator buffer

    if (warp == 0) {
        if (elect_one()) {
            pdl_wait();
            int it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int w0 = (t % p.tiles_w) * 8, h0 = ((t / p.tiles_w) % p.tiles_h) * 16, n0 = t / (p.tiles_w * p.tiles_h);
                for (int dxi = 0; dxi < 3; ++dxi, ++it) {
                    const int s = it % S;
                    if (it >= S) mbar_wait(&a_empty[s], (uint32_t)((it / S) - 1) & 1u, 1);
                    mbar_expect_tx(&a_full[s], (uint32_t)kHaloABytes);
                    tma_load_4d(sA + (size_t)s * kHaloABytes, &mapA, &a_full[s], 0, w0 + dxi - 1, h0 - 1, n0);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_bf16(kBlockM, (uint32_t)p.block_n);
        mbar_wait(b_full, 0, 2);
        int it = 0, i = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
            const int ab = i & 1, use = i >> 1;
            if (use > 0) mbar_wait(&acc_empty[ab], (uint32_t)(use - 1) & 1u, 3);   // the epilogue has drained this buffer
            tc_fence_after_sync();
            const uint32_t tacc = tmem_base + (uint32_t)ab * acc_cols;
            for (int dxi = 0; dxi < 3; ++dxi, ++it) {
                const int s = it % S;
                mbar_wait(&a_full[s], (uint32_t)(it / S) & 1u, 4);
                tc_fence_after_sync();
                const uint32_t a_addr = smem_u32(sA + (size_t)s * kHaloABytes);
                if (elect_one()) {
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        const uint32_t b_addr = smem_u32(sB + (size_t)(dyi * 3 + dxi) * b_bytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16(tacc, umma_desc_sw128(a_addr + dyi * 1024 + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                      (dxi == 0 && dyi == 0 && k == 0) ? 0u : 1u);
                    }
                    umma_commit(&a_empty[s]);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(&acc_full[ab]);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int bw = r & 7, bh = r >> 3;                      // 8 x 16 tile
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        pdl_wait();
        int i = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++i) {
            const int ab = i & 1, use = i >> 1;
            const int w0 = (t % p.tiles_w) * 8, h0 = ((t / p.tiles_w) % p.tiles_h) * 16, n0 = t / (p.tiles_w * p.tiles_h);
            const int hh = h0 + bh, ww = w0 + bw;
            const bool row_ok = (hh < p.H) && (ww < p.W);
            const long grow = ((long)n0 * p.H + hh) * p.W + ww;
            // the residual row of this thread is requested before the accumulator is waited for
            uint4 rres[8];
            const bool res_pre = (p.residual != nullptr) && row_ok && ((p.ldr & 7) == 0) && (p.block_n <= 64);
            if (res_pre) {
                const uint4* r4 = reinterpret_cast<const uint4*>(p.residual + grow * p.ldr);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j * 8 < p.N) rres[j] = r4[j];
            }
            mbar_wait(&acc_full[ab], (uint32_t)use & 1u, 5);
            tc_fence_after_sync();
            const uint32_t tacc = tmem_base + lane_off + (uint32_t)ab * acc_cols;
            // the previous tile's bulk stores must have read the staging buffers before they are overwritten
            tma_store_wait_all();
            __syncwarp();
            for (int c = 0; c < p.block_n; c += 32) {
                uint32_t u[32];
                tmem_ld32(tacc + c, u);
                tmem_ld_wait();
                if (c + 32 >= p.block_n) {                      // last read of this accumulator: hand it back to the MMA warp
                    tc_fence_before_sync();
                    mbar_arrive(&acc_empty[ab]);
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]);
                const int ncols = min(32, p.N - c);
                if (row_ok) {
                    if (res_pre && ncols == 32) {
                        uint4 rc[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) rc[j] = (c == 0) ? rres[j] : rres[4 + j];
                        epilogue_math32(p, v, n0, grow, c, ncols, sbias + c, false, rc);
                    } else {
                        epilogue_math32(p, v, n0, grow, c, ncols, sbias + c, false, nullptr);
                    }
                }
                uint8_t* buf = stag + (size_t)((c >> 5) * 4 + q) * 2048;
                uint8_t* myrow = buf + lane * 64;
                const int sx = (lane >> 1) & 3;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(myrow + ((j ^ sx) << 4)) =
                        make_uint4(pack_bf16x2(v[j * 8], v[j * 8 + 1]), pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]),
                                   pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]));
                fence_proxy_async_smem();
                __syncwarp();
                if (ncols > 0 && elect_one()) {
                    tma_store_4d(&mapC, buf, c, w0, h0 + q * 4, n0);   // warp q: rows 32q .. 32q+31 = image rows h0+4q .. +3
                    tma_store_commit();
                }
            }
        }
        tma_store_wait_all();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 2, 1u);
}

// Sums split-K partials (fixed order => deterministic) and applies the epilogue. One thread per (row, 4 columns):
// many threads with one float4 per split each keep plenty of loads in flight for this latency-bound pass.
__global__ void splitk_reduce_kernel(const GemmParams p, long rows) {
    pdl_launch_dependents();
    pdl_wait();
    const int groups4 = (p.N + 3) >> 2;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * groups4) return;
    const long grow = idx / groups4;
    const int col = (int)(idx - grow * groups4) * 4;
    const int ncols = min(4, p.N - col);
    const bool vec = (ncols == 4) && ((p.N & 3) == 0);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float* src = p.partial + grow * p.N + col;
    const long split_stride = rows * p.N;
    if (vec) {
#pragma unroll 8
        for (int s = 0; s < p.splits; ++s) {
            const float4 t = *reinterpret_cast<const float4*>(src + (long)s * split_stride);
            v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
        }
    } else {
        for (int s = 0; s < p.splits; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < ncols) v[j] += src[(long)s * split_stride + j];
    }
    splitk_finish4(p, v, grow, col);
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_encode_once;

static void resolve_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, const cuuint32_t* elem_strides = nullptr) {
    std::call_once(g_encode_once, resolve_encode);
    VSD_REQUIRE(g_encode != nullptr, "cuTensorMapEncodeTiled driver entry point not available (no CUDA driver?)");
    VSD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base must be 16-byte aligned");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (elem_strides)
        for (int i = 0; i < rank; ++i) estr[i] = elem_strides[i];
    CUresult r = g_encode(m, dtype, (cuuint32_t)rank, const_cast<void*>(base), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " rank=" +
                  std::to_string(rank) + " dims=" + std::to_string(dims[0]) + "," + std::to_string(dims[1]) +
                  " box=" + std::to_string(box[0]) + "," + std::to_string(box[1]));
        return -3;
    }
    return 0;
}

// elem_stride 2: the box spans boxW * 2 x boxH * 2 input pixels and TMA delivers every second one (stride-2 convolution taps)
int make_tmap_act(CUtensorMap* m, const void* base, int C, int W, int H, int N, int ld, int boxW, int boxH, int boxN,
                  int elem_stride) {
    VSD_REQUIRE((ld % 8) == 0, "activation pixel stride must be a multiple of 8 elements");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(boxW * elem_stride), (cuuint32_t)(boxH * elem_stride), (cuuint32_t)boxN};
    cuuint32_t es[4] = {1, (cuuint32_t)elem_stride, (cuuint32_t)elem_stride, 1};
    return encode(m, base, 4, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, es);
}

// Epilogue maps: 32-column boxes over a [NB][H][W][ld] bf16 tensor (64-byte swizzle), or over the fp32 split-K workspace
// [splits][NB][H][W][N] (128-byte swizzle).
static int make_tmap_epi_bf16(CUtensorMap* m, const void* base, int C, int W, int H, int N, int ld, int boxW, int boxH, int boxN) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxN};
    return encode(m, base, 4, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B);
}
static int make_tmap_epi_partial(CUtensorMap* m, const void* base, int C, int W, int H, int N, int splits, int boxW, int boxH, int boxN) {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)splits};
    cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)N * H * W * C * 4};
    cuuint32_t box[5] = {32, (cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxN, 1};
    return encode(m, base, 5, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_2d(CUtensorMap* m, const void* base, int K, int rows, int ld, int box_rows) {
    VSD_REQUIRE((ld % 8) == 0, "matrix row stride must be a multiple of 8 elements");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    return encode(m, base, 2, dims, strides, box);
}

static int g_num_sms = 148;
static int g_max_smem = 227 * 1024;

// The instantiated (epilogue kind, pairs, feature set) combinations; a launch needing `feat` gets the smallest instantiated
// superset (*got). Epilogue 0 (fp32 / unaligned outputs: rare) and any unusual combination run the all-features kernel.
typedef void (*GemmKernel)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, GemmParams);
static GemmKernel gemm_kernel_for(int epi, int pair, int feat, int* got) {
#define VSD_K(E, P, F) do { *got = (F); return conv_gemm_kernel<E, P, F>; } while (0)
    if (epi == 0) { if (pair) VSD_K(0, true, FEAT_ALL); VSD_K(0, false, FEAT_ALL); }
    if (epi == 2) { if (pair) VSD_K(2, true, 0); VSD_K(2, false, 0); }
    if (epi == 3) { if (pair) return nullptr; if (feat & FEAT_STATS) VSD_K(3, false, FEAT_STATS); VSD_K(3, false, 0); }
    if (epi == 1 && pair) {
        switch (feat) {
            case 0: VSD_K(1, true, 0);
            case FEAT_GEGLU: VSD_K(1, true, FEAT_GEGLU);
            case FEAT_LN: VSD_K(1, true, FEAT_LN);
            case FEAT_GEGLU | FEAT_LN: VSD_K(1, true, FEAT_GEGLU | FEAT_LN);
            case FEAT_STATS: VSD_K(1, true, FEAT_STATS);
            default: VSD_K(1, true, FEAT_ALL);
        }
    }
    if (epi == 1) {
        switch (feat) {
            case 0: VSD_K(1, false, 0);
            case FEAT_GEGLU: VSD_K(1, false, FEAT_GEGLU);
            case FEAT_LN: VSD_K(1, false, FEAT_LN);
            case FEAT_GEGLU | FEAT_LN: VSD_K(1, false, FEAT_GEGLU | FEAT_LN);
            case FEAT_STATS: VSD_K(1, false, FEAT_STATS);
            default: VSD_K(1, false, FEAT_ALL);
        }
    }
#undef VSD_K
    return nullptr;
}

int gemm_init() {
    int dev = 0;
    VSD_CHECK_CUDA(cudaGetDevice(&dev));
    VSD_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    VSD_CHECK_CUDA(cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    for (int epi = 0; epi < 4; ++epi)
        for (int pair = 0; pair < 2; ++pair)
            for (int feat = 0; feat <= FEAT_ALL; ++feat) {
                int got = -1;
                GemmKernel k = gemm_kernel_for(epi, pair, feat, &got);
                if (k && got == feat)   // each instantiation once
                    VSD_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
            }
    VSD_CHECK_CUDA(cudaFuncSetAttribute(conv_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_max_smem));
    return 0;
}

static int pow2_at_least(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

// Pick the pixel rectangle (BW x BH x BN, powers of two, product 128) covered by one 128-row tile: the shape
// that needs the fewest tiles; ties go to the widest rectangle (longest contiguous TMA rows).
static void pick_tile_rect(int NB, int H, int W, int* BW, int* BH, int* BN) {
    long best_tiles = -1;
    int bbw = 128, bbh = 1, bbn = 1;
    for (int bw = 128; bw >= 1; bw >>= 1) {
        for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
            const int bn = 128 / (bw * bh);
            const long tiles = (long)((W + bw - 1) / bw) * ((H + bh - 1) / bh) * ((NB + bn - 1) / bn);
            if (best_tiles < 0 || tiles < best_tiles) {
                best_tiles = tiles; bbw = bw; bbh = bh; bbn = bn;
            }
        }
    }
    *BW = bbw; *BH = bbh; *BN = bbn;
}

// Persistent weight-stationary 3x3 convolution (conv_persist_kernel): 64 input channels, <= 64 output channels, bf16 output.
static int build_persist_op(GemmOp* op, const ActView& a, const bf16* wt, int N, int ldw, void* out, int ldo, int out_f32,
                            const float* bias, const float* rowvec, const bf16* residual, int ldr, int act_flags) {
    GemmParams& p = op->p;
    VSD_REQUIRE(p.taps == 9 && a.C == 64 && N >= 8 && N <= 64 && a.H >= 16 && a.W >= 8, "persistent conv: 3x3, 64 -> <= 64 channels");
    VSD_REQUIRE(!out_f32 && rowvec == nullptr && (ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (N % 8) == 0 &&
                (act_flags & 0xF) == ACT_NONE && !(act_flags & ACT_RES_F32_FLAG), "persistent conv: bf16 output, no row vector");
    p.persist = 1; p.halo = 1; p.pair = 0;
    p.BW = 8; p.BH = 16; p.BN = 1;
    p.tiles_w = (a.W + 7) / 8; p.tiles_h = (a.H + 15) / 16; p.tiles_n = a.NB;
    p.N = N;
    p.block_n = N <= 32 ? 32 : 64;
    p.tmem_cols = 2 * p.block_n;                      // two accumulator buffers (64 or 128 columns, powers of two)
    p.splits = 1; p.kb_total = 3; p.kb_per_split = 3; p.kb_per_stage = 1;
    p.stages = 4;
    p.out = out; p.ldo = ldo; p.out_f32 = 0;
    p.bias = bias; p.rowvec = nullptr; p.residual = residual; p.ldr = ldr; p.res_f32 = 0;
    p.act = ACT_NONE; p.relu = (act_flags & ACT_RELU_FLAG) ? 1 : 0;
    p.tma_out = 1; p.tma_res = 0; p.cluster_k = 0;
    p.sbw = 8; p.sbh = 4; p.sbn = 1;                  // a warp's 32 rows = 8 x 4 pixels
    p.lbw = 3; p.lbh = 4;
    const int region = 9 * p.block_n * 128 + p.stages * kHaloABytes + p.block_n * 256;
    p.stage_off = 0;
    p.bar_off = (unsigned)region;
    op->smem_bytes = region + 1024 + (1 + 2 * p.stages + 4) * 8 + 64 + p.block_n * 4 + 64;
    VSD_REQUIRE(op->smem_bytes <= g_max_smem, "persistent conv does not fit shared memory");
    int rc = make_tmap_act(&op->mapA, a.ptr, a.C, a.W, a.H, a.NB, a.ld, 8, 18, 1);
    if (rc) return rc;
    rc = make_tmap_2d(&op->mapB, wt, 9 * a.C, N, ldw, p.block_n);
    if (rc) return rc;
    rc = make_tmap_epi_bf16(&op->mapC, out, N, a.W, a.H, a.NB, ldo, 8, 4, 1);
    if (rc) return rc;
    op->mapR = op->mapA;
    const int tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    op->grid = dim3(tiles < g_num_sms ? tiles : g_num_sms, 1, 1);
    return 0;
}

int build_gemm_op(GemmOp* op, const ActView& ain, int taps, const bf16* wt, int N, int ldw, void* out, int ldo,
                  int out_f32, const float* bias, const float* rowvec, const bf16* residual, int ldr, int act_flags,
                  float* partial_ws, size_t partial_ws_bytes, int force_block_n, int force_splits, int force_occupancy,
                  int force_kb_per_stage, int force_halo, const LnFuse* ln) {
    // `a` is the OUTPUT geometry from here on (tiles, rows, epilogue maps); `ain` only describes the activation tensor map
    const int cs = (taps == 9 && ain.stride == 2) ? 2 : 1;
    VSD_REQUIRE(ain.stride == 1 || (ain.stride == 2 && taps == 9), "only 3x3 convolutions can be strided (stride 2)");
    ActView a = ain;
    if (cs == 2) {
        a.H = (ain.H + 2 * ain.pad - 3 + (ain.pad ? 0 : 1)) / 2 + 1;   // pad 1: (H-1)/2+1 ; pad 0 with one zero row below: (H-2)/2+1
        a.W = (ain.W + 2 * ain.pad - 3 + (ain.pad ? 0 : 1)) / 2 + 1;
        if (force_halo > 0) force_halo &= ~5;        // halo tiles / the persistent kernel assume stride 1
        VSD_REQUIRE(a.H >= 1 && a.W >= 1, "strided convolution output is empty");
    }
    const int act = act_flags & 0xF;
    VSD_REQUIRE(taps == 1 || taps == 9, "taps must be 1 or 9");
    VSD_REQUIRE(a.C % 64 == 0, "input channels must be a multiple of 64 for the tcgen05 path");
    VSD_REQUIRE(a.ld >= a.C, "bad activation stride");
    GemmParams& p = op->p;
    p = GemmParams{};
    p.taps = taps; p.cin = a.C; p.H = a.H; p.W = a.W; p.NB = a.NB;
    // Halo mode for 3x3 convolutions: 8 x 16 pixel tiles; per 64-channel block three column-shifted 8 x 18 halo
    // tiles replace nine 128-pixel tap tiles (2.7x less activation traffic into the SM).
    // force_halo is a mode word: bit 0 = halo tiles, bit 1 = CTA pairs (cta_group::2)
    const bool halo = (taps == 9) && (a.H >= 16) && (a.W >= 8) && (force_halo > 0) && (force_halo & 1);
    const bool pair = (force_halo > 0) && (force_halo & 2);
    const bool persist = (force_halo > 0) && (force_halo & 4);
    p.halo = halo ? 1 : 0;
    if (halo) { p.BW = 8; p.BH = 16; p.BN = 1; }
    else pick_tile_rect(a.NB, a.H, a.W, &p.BW, &p.BH, &p.BN);
    p.tiles_w = (a.W + p.BW - 1) / p.BW;
    p.tiles_h = (a.H + p.BH - 1) / p.BH;
    p.tiles_n = (a.NB + p.BN - 1) / p.BN;
    p.N = N;
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    p.pair = pair ? 1 : 0;
    p.cstride = cs;
    p.cshift = (cs == 2 && ain.pad == 0) ? 1 : 0;
    p.persist = 0;
    if (ln && ln->mode) {
        VSD_REQUIRE(taps == 1 && !persist && cs == 1, "LayerNorm fold: linear layers only");
        VSD_REQUIRE(force_splits <= 1, "LayerNorm fold: split-K not supported (the epilogue is not linear in the partial sums)");
        VSD_REQUIRE(ln->wsum != nullptr && ln->stats != nullptr && ln->nst >= 1 && (ln->mode == 1 || ln->mode == 2),
                    "LayerNorm fold: missing column sums / row statistics");
        force_splits = 1;
    }
    if (ln && ln->stats_out) VSD_REQUIRE(!persist && (act_flags & 0xF) == ACT_NONE, "row statistics: plain epilogue only");
    if (persist) return build_persist_op(op, a, wt, N, ldw, out, ldo, out_f32, bias, rowvec, residual, ldr, act_flags);
    p.kb_total = halo ? 3 * (a.C / 64) : taps * (a.C / 64);   // halo: iterations of (channel block, column shift)

    // N tile: prefer a divisor of N that keeps the grid near a multiple of the SM count.
    int bn = 0;
    if (force_block_n > 0) {
        bn = force_block_n;
    } else if (act == ACT_GEGLU) {
        bn = 128;
    } else {
        const int cands[6] = {256, 160, 128, 96, 64, 32};
        double best_score = -1;
        for (int i = 0; i < 6; ++i) {
            const int c = cands[i];
            if (c > ((N + 31) / 32) * 32 && c != 32) continue;
            const int n_tiles = (N + c - 1) / c;
            const double waste = (double)(n_tiles * c) / (double)N;      // padded columns
            const long ctas = (long)m_tiles * n_tiles;
            const double waves = (double)ctas / g_num_sms;
            const double eff = waves / (double)((long)(waves + 0.999999));  // tail efficiency
            // bigger tiles amortise A smem traffic; small tiles fill the machine
            const double tile_eff = (c >= 128) ? 1.0 : (c >= 96 ? 0.9 : (c >= 64 ? 0.8 : 0.6));
            const double score = eff * tile_eff / waste;
            if (score > best_score) { best_score = score; bn = c; }
        }
    }
    VSD_REQUIRE(bn % 32 == 0 && bn >= 32 && bn <= 256, "block_n must be a multiple of 32 in [32,256]");
    if (act == ACT_GEGLU) VSD_REQUIRE(bn % 128 == 0 && N % bn == 0 && bias != nullptr, "GEGLU needs block_n = 128 or 256 dividing N, and a bias");
    p.block_n = bn;
    p.tmem_cols = pow2_at_least(bn);
    const int n_tiles = (N + bn - 1) / bn;

    // split K when the tile grid cannot fill the machine and K is deep
    int splits = 1;
    if (force_splits > 0) {
        splits = force_splits;
    } else if (act == ACT_NONE && partial_ws != nullptr) {
        const long ctas = (long)m_tiles * n_tiles;
        if (ctas * 2 <= g_num_sms && p.kb_total >= 16) {
            splits = (int)(g_num_sms / ctas);
            const int max_by_k = p.kb_total / 8;
            if (splits > max_by_k) splits = max_by_k;
            if (splits > 16) splits = 16;
            if (splits < 1) splits = 1;
        }
    }
    const long rows = (long)a.NB * a.H * a.W;
    if (force_splits <= 0)
        while (splits > 1 && (size_t)splits * rows * N * 4 > partial_ws_bytes) --splits;  // fit the workspace
    p.kb_per_split = (p.kb_total + splits - 1) / splits;
    splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.splits = splits;
    if (splits > 1) {
        VSD_REQUIRE(partial_ws != nullptr && (size_t)splits * rows * N * 4 <= partial_ws_bytes,
                    "split-K workspace too small");   // (the cluster variant does not touch it; kept as the common bound)
        VSD_REQUIRE(act == ACT_NONE, "split-K cannot be combined with GEGLU");
    }
    p.partial = partial_ws;

    // Epilogue through shared memory + bulk tensor stores (see GemmParams::tma_out)
    static const bool tma_epi = !(getenv("VSD_TMA_EPI") && atoi(getenv("VSD_TMA_EPI")) == 0);
    const int n_out = (act == ACT_GEGLU) ? N / 2 : N;
    const int bn_out = (act == ACT_GEGLU) ? bn / 2 : bn;
    int tma_out = 0, tma_res = 0, cluster_k = 0;
    // Split-K inside a thread-block cluster (1,1,splits): accumulators are exchanged through distributed shared memory and
    // reduced by the cluster itself -- no fp32 partials through L2, no second kernel (conv_gemm_kernel<3>). Mode word bit 3
    // requests it, bit 4 forbids it (the autotuner times both); without either it is used whenever it applies.
    static const bool cluster_default = !(getenv("VSD_CLUSTER_SPLITK") && atoi(getenv("VSD_CLUSTER_SPLITK")) == 0);
    const bool want_cluster = (force_halo > 0 && (force_halo & 24)) ? ((force_halo & 8) != 0) : cluster_default;
    const bool out_tma_ok = !out_f32 && (ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (bn_out % 32) == 0;
    if (splits > 1 && splits <= 8 && want_cluster && !pair && (N % 4) == 0) {
        cluster_k = 1;
    } else if (tma_epi) {
        if (splits > 1) {
            if ((N % 4) == 0 && (reinterpret_cast<uintptr_t>(partial_ws) & 15) == 0) tma_out = 2;
        } else if (out_tma_ok) {
            tma_out = 1;
            if (residual != nullptr && !(act_flags & ACT_RES_F32_FLAG) && (ldr % 8) == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0) tma_res = 1;
        }
    }
    int stag_bytes = (tma_out == 2) ? bn * 512 : (tma_out == 1 ? bn_out * 256 : 0);   // 128 rows x (4 | 2) bytes per column
    if (cluster_k) stag_bytes = splits * ((kBlockM + splits - 1) / splits) * bn * 4;   // receive buffer: `splits` copies of this CTA's row band

    // Tiles up to 160 columns run two CTAs per SM (one CTA's epilogue overlaps the other's main loop);
    // wider tiles take the whole SM with a deeper ring.
    const int b_tile = pair ? bn * 64 : bn * 128;     // bytes of one k-block of the weight tile held by one CTA
    const int stage_bytes = kABytes + b_tile;
    int occ = force_occupancy > 0 ? force_occupancy : ((bn > 160) ? 1 : 2);
    if (pair) occ = 1;
    if (occ == 2 && 2 * (kABytes + b_tile) + 4096 > g_max_smem / 2) occ = 1;   // not even two stages fit twice
    if (halo && occ == 2 && force_occupancy <= 0 && 2 * (kHaloABytes + 3 * b_tile) + 4096 > g_max_smem / 2) occ = 1;
    const int smem_budget = (occ == 1) ? g_max_smem : (g_max_smem / 2 - 1024);
    // Each barrier round trip (TMA -> full -> MMA -> commit -> empty -> TMA) costs several hundred cycles, so a
    // stage carries kb_per_stage 64-wide k-blocks; keep >= 3 stages in flight when the budget allows.
    int kbs = force_kb_per_stage > 0 ? force_kb_per_stage : 2;
    if (halo) kbs = 1;
    int stages = 0, stage_total = 0;
    for (;;) {
        // the residual tile needs its own region (it is in flight while the ring is busy); plain staging reuses the ring
        // reserve: alignment slack, barriers, staged bias (+ the folded-LayerNorm column vectors)
        const int ring_budget = smem_budget - 3072 - ((ln && ln->mode) ? 2 * bn * 4 : 0) - (tma_res ? stag_bytes : 0);
        int k2 = kbs;
        while (k2 > 1 && (ring_budget / (k2 * stage_bytes) < (force_kb_per_stage > 0 ? 2 : 3) || k2 > p.kb_per_split)) --k2;
        stage_total = halo ? (kHaloABytes + 3 * b_tile) : k2 * stage_bytes;
        stages = ring_budget / stage_total;
        if (tma_res && stages < 2) { tma_res = 0; continue; }   // no room: read the residual straight from global memory
        if (cluster_k && stag_bytes > smem_budget - 3072) { cluster_k = 0; stag_bytes = 0; if (tma_epi && (reinterpret_cast<uintptr_t>(partial_ws) & 15) == 0) { tma_out = 2; stag_bytes = bn * 512; } continue; }
        if (tma_out && !tma_res && stag_bytes > smem_budget - 3072) { tma_out = 0; stag_bytes = 0; continue; }
        kbs = k2;
        break;
    }
    if (halo) VSD_REQUIRE(stages >= 2, "halo tile does not leave room for two pipeline stages");
    if (stages > 8) stages = 8;
    const int stage_iters = (p.kb_per_split + kbs - 1) / kbs;
    if (stages > stage_iters) stages = stage_iters;
    if (stages < 1) stages = 1;
    p.stages = stages;
    p.kb_per_stage = kbs;
    const int ring_bytes = stages * stage_total;
    int region = ring_bytes;
    p.stage_off = 0;
    if (tma_res) { p.stage_off = (unsigned)ring_bytes; region = ring_bytes + stag_bytes; }
    else if (stag_bytes > region) region = stag_bytes;
    p.bar_off = (unsigned)region;
    p.tma_out = tma_out; p.tma_res = tma_res; p.cluster_k = cluster_k;
    p.lbw = 0; while ((1 << p.lbw) < p.BW) ++p.lbw;
    p.lbh = 0; while ((1 << p.lbh) < p.BH) ++p.lbh;
    p.sbw = p.BW < 32 ? p.BW : 32;
    p.sbh = (32 / p.sbw) < p.BH ? (32 / p.sbw) : p.BH;
    p.sbn = 32 / (p.sbw * p.sbh);
    op->smem_bytes = region + 1024 /*align slack*/ + (2 * stages + 2) * 8 + 64 + bn * 4;
    p.ln_mode = 0; p.ln_wsum = nullptr; p.ln_rowbias = nullptr; p.ln_eps = 0.f; p.ln_stats = nullptr; p.ln_nst = 0;
    p.rowstats_out = nullptr;
    if (ln && ln->mode) {
        VSD_REQUIRE(tma_out == 1 && splits == 1, "LayerNorm fold needs the bf16 tensor-store epilogue");
        p.ln_mode = ln->mode; p.ln_wsum = ln->wsum; p.ln_rowbias = ln->rowbias; p.ln_eps = ln->eps;
        p.ln_stats = ln->stats; p.ln_nst = ln->nst;
        op->smem_bytes += 2 * bn * 4;   // column sums / column statistics next to the staged bias
    }
    if (ln && ln->stats_out) {
        // [rows][n_tiles][2] partial sums written by the epilogue that stores the rows: the bf16 tensor-store epilogue or the
        // in-cluster split-K reduction (the separate split-K reduce kernel has no thread that sees a whole row)
        VSD_REQUIRE((tma_out == 1 && splits == 1) || cluster_k, "row statistics need the tensor-store epilogue or the in-cluster split-K reduction");
        p.rowstats_out = ln->stats_out;
    }
    // With programmatic dependent launch CTAs of different kernels co-reside on an SM. TMEM is not part of the block
    // scheduler's accounting, so bound the CTAs per SM through shared memory: smem >= tmem_cols * 450 B guarantees that
    // the co-resident CTAs' TMEM columns sum to <= 512 (otherwise tcgen05.alloc of a CTA the others wait on could spin).
    if (op->smem_bytes < p.tmem_cols * 450) op->smem_bytes = p.tmem_cols * 450;
    // Capacity is not contiguity: with kernels of several streams (lanes, the ControlNet branch) on one SM a tcgen05.alloc may
    // wait for a neighbour to finish. That is harmless for a CTA nobody waits on, but a cluster CTA (CTA pair, in-cluster
    // split-K) is waited on by its peers *while they hold their own columns*: cluster {a on SM0, b on SM1} and cluster
    // {c on SM1, d on SM0} with a, c allocated and b, d waiting for space a, c fragment is a cycle. Cluster kernels therefore
    // take their SMs alone (no second TMEM user fits beside them: the smallest needs 32 * 450 B + 1 KB), so their allocation
    // never waits on a running CTA and every wait chain ends at a kernel that needs nothing more.
    static const int cluster_alone = getenv("VSD_CLUSTER_ALONE") ? atoi(getenv("VSD_CLUSTER_ALONE")) : 1;
    if ((pair || cluster_k) && cluster_alone && op->smem_bytes < g_max_smem - 8192) op->smem_bytes = g_max_smem - 8192;
    VSD_REQUIRE(op->smem_bytes <= g_max_smem, "GEMM configuration does not fit shared memory");

    p.out = out; p.ldo = ldo; p.out_f32 = out_f32;
    p.bias = bias; p.rowvec = rowvec; p.residual = residual; p.ldr = ldr;
    p.res_f32 = (act_flags & ACT_RES_F32_FLAG) ? 1 : 0;
    p.act = act;
    p.relu = (act_flags & ACT_RELU_FLAG) ? 1 : 0;
    p.a_static = (act_flags & ACT_A_STATIC_FLAG) ? 1 : 0;
    p.b_static = (act_flags & (ACT_A_STATIC_FLAG | ACT_NO_STATIC_FLAG)) ? 0 : 1;
    if (p.a_static && taps != 1) p.a_static = 0;

    int rc = halo ? make_tmap_act(&op->mapA, a.ptr, a.C, a.W, a.H, a.NB, a.ld, 8, 18, 1)
                  : make_tmap_act(&op->mapA, ain.ptr, ain.C, ain.W, ain.H, ain.NB, ain.ld, p.BW, p.BH, p.BN, cs);
    if (rc) return rc;
    rc = make_tmap_2d(&op->mapB, wt, taps * a.C, N, ldw, pair ? bn / 2 : bn);
    if (rc) return rc;
    op->mapC = op->mapA;   // placeholders when unused (a valid descriptor keeps the kernel parameter well-formed)
    op->mapR = op->mapA;
    if (tma_out == 1) rc = make_tmap_epi_bf16(&op->mapC, out, n_out, a.W, a.H, a.NB, ldo, p.sbw, p.sbh, p.sbn);
    else if (tma_out == 2) rc = make_tmap_epi_partial(&op->mapC, partial_ws, N, a.W, a.H, a.NB, splits, p.sbw, p.sbh, p.sbn);
    if (rc) return rc;
    if (tma_res) rc = make_tmap_epi_bf16(&op->mapR, residual, N, a.W, a.H, a.NB, ldr, p.BW, p.BH, p.BN);
    if (rc) return rc;
    // pairs: adjacent 128-row tiles share one cluster; an odd tile count is padded with a tile outside the tensor (TMA reads
    // zeros / clips the stores, the direct epilogue masks the rows)
    op->grid = dim3(pair ? (m_tiles + 1) & ~1 : m_tiles, n_tiles, splits);
    return 0;
}

// Every (block_n, split-K, CTAs/SM, k-blocks per stage, mode) configuration of the kernel family that is valid for a shape:
// the list the engine's autotuner times (vsd_engine.cu: tune_gemm) and the operator-level sweep test checks against fp32
// (tests/test_gpu_tuner_sweep.py), so a configuration the tuner may pick is a configuration that was tested.
// mode word: bit 0 halo tiles, bit 1 CTA pairs, bit 2 persistent weight-stationary kernel, bit 3 split-K reduced inside the
// cluster, bit 4 separate reduce kernel.
int enumerate_gemm_candidates(const ActView& a, int taps, const bf16* wt, int N, int ldw, void* outp, int ldo, int out_f32,
                              const float* bias, const float* rowvec, const bf16* res, int ldr, int act, float* ws, size_t ws_bytes,
                              const LnFuse* ln, std::vector<GemmCand>* out) {
    out->clear();
    const bool geglu = (act & 0xF) == ACT_GEGLU;
    const bool ln_consumer = ln != nullptr && ln->mode != 0;   // no split-K (the epilogue is not linear in the partial sums)
    const int bns[8] = {32, 64, 96, 128, 160, 192, 224, 256};
    const int kbss[3] = {1, 2, 4};
    const int sps[8] = {1, 2, 3, 4, 6, 8, 12, 16};
    const int kb_total = taps * (a.C / 64);
    static const int persist_ok = !(getenv("VSD_TUNE_PERSIST") && atoi(getenv("VSD_TUNE_PERSIST")) == 0);
    static const int pairs_ok = !(getenv("VSD_TUNE_PAIRS") && atoi(getenv("VSD_TUNE_PAIRS")) == 0);
    static const int only_occ = getenv("VSD_TUNE_OCC") ? atoi(getenv("VSD_TUNE_OCC")) : 0;   // experiment knob
    GemmOp op;
    // persistent weight-stationary variant for the TAESD-shaped 3x3 convolutions (mode 4)
    if (persist_ok && taps == 9 && a.stride == 1 && a.C == 64 && N <= 64 && N % 8 == 0 && !out_f32 && rowvec == nullptr && a.H >= 16 &&
        a.W >= 8 && (act & 0xF) == ACT_NONE && ln == nullptr &&
        !build_gemm_op(&op, a, taps, wt, N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act, ws, ws_bytes, 0, 1, 1, 1, 4))
        out->push_back(GemmCand{op.p.block_n, 1, 1, 1, 4});
    for (int bi = 0; bi < 8; ++bi) {
        const int bn = bns[bi];
        if (geglu && bn != 128 && bn != 256) continue;
        if (bn != 32 && bn > ((N + 31) / 32) * 32) continue;
        for (int si = 0; si < 8; ++si) {
            const int sp = sps[si];
            if (sp > 1 && (geglu || ln_consumer || kb_total / sp < 2)) continue;
            for (int occ = 1; occ <= 2; ++occ) {
                if (only_occ && occ != only_occ) continue;
                for (int ki = 0; ki < 4; ++ki) {
                    const int use_halo = (ki == 3) ? 1 : 0;             // 4th variant: 3x3 halo mode
                    if (use_halo && (taps != 9 || a.stride != 1 || a.H < 16 || a.W < 8)) continue;
                    const int kbs = use_halo ? 1 : kbss[ki];
                    for (int pc = 0; pc < 3; ++pc) {                    // plain | CTA pairs (cta_group::2, whole-SM CTAs only) | in-cluster split-K
                        const int pair = pc == 1 ? 1 : 0, ck = pc == 2 ? 1 : 0;
                        if (pair && (occ == 2 || !pairs_ok || (ln_consumer && ln->mode == 2))) continue;   // swapped-operand LN: no pairs
                        if (ck && (sp < 2 || sp > 8)) continue;
                        const int mode = use_halo | (pair << 1) | (ck ? 8 : 16);
                        if (build_gemm_op(&op, a, taps, wt, N, ldw, outp, ldo, out_f32, bias, rowvec, res, ldr, act, ws, ws_bytes, bn, sp, occ,
                                          kbs, mode, ln))
                            continue;   // does not fit (workspace / smem) or not offered for this epilogue: skip
                        if (op.p.splits != sp || op.p.kb_per_stage != kbs || op.p.halo != use_halo || op.p.pair != pair || op.p.cluster_k != ck) continue;
                        const long ctas = (long)op.grid.x * op.grid.y * op.grid.z;
                        if (sp > 1 && ctas > 4 * 148) continue;
                        if (occ == 2 && op.smem_bytes > 114 * 1024) continue;   // would not actually co-reside
                        out->push_back(GemmCand{bn, sp, occ, kbs, mode});
                    }
                }
            }
        }
    }
    return (int)out->size();
}

int launch_gemm_op(const GemmOp& op, cudaStream_t st) {
    if (op.p.persist) {
        VSD_CHECK_CUDA(launch_k(conv_persist_kernel, op.grid, dim3(kGemmThreads), (size_t)op.smem_bytes, st, op.mapA, op.mapB, op.mapC, op.p));
        return 0;
    }
    const GemmParams& q = op.p;
    const int feat = ((q.act == ACT_GEGLU) ? FEAT_GEGLU : 0) | (q.ln_mode ? FEAT_LN : 0) | (q.rowstats_out ? FEAT_STATS : 0) |
                     ((q.act == ACT_QUICK_GELU || q.out_scale != nullptr || q.res_f32) ? FEAT_MISC : 0);
    int got = 0;
    if (op.p.cluster_k) {
        GemmKernel kern = gemm_kernel_for(3, 0, feat, &got);
        VSD_CHECK_CUDA(launch_k_cluster(kern, op.grid, dim3(kGemmThreads), (size_t)op.smem_bytes, 1, op.p.splits, st,
                                        op.mapA, op.mapB, op.mapC, op.mapR, op.p));
        return 0;
    }
    GemmKernel kern = gemm_kernel_for(op.p.tma_out, op.p.pair ? 1 : 0, feat, &got);
    VSD_REQUIRE(kern != nullptr, "no GEMM kernel instantiation for this configuration");
    if (op.p.pair) {
        VSD_CHECK_CUDA(launch_k_cluster(kern, op.grid, dim3(kGemmThreads), (size_t)op.smem_bytes, 2, 1, st, op.mapA, op.mapB, op.mapC,
                                        op.mapR, op.p));
    } else {
        VSD_CHECK_CUDA(launch_k(kern, op.grid, dim3(kGemmThreads), (size_t)op.smem_bytes, st, op.mapA, op.mapB, op.mapC, op.mapR, op.p));
    }
    if (op.p.splits > 1) {
        const long rows = (long)op.p.NB * op.p.H * op.p.W;
        return launch_splitk_reduce(op.p, rows, st);
    }
    return 0;
}

int launch_splitk_reduce(const GemmParams& p, long rows, cudaStream_t st) {
    const long work = rows * ((p.N + 3) / 4);
    const int threads = 256;
    VSD_CHECK_CUDA(launch_k(splitk_reduce_kernel, dim3((unsigned)((work + threads - 1) / threads)), dim3(threads), 0, st, p, rows));
    return 0;
}

const unsigned int* trap_code_addr_gemm() {
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_trap_code) == cudaSuccess ? static_cast<const unsigned int*>(p) : nullptr;
}

unsigned int read_trap_code_gemm() {
    unsigned int v = 0, z = 0;
    if (cudaMemcpyFromSymbol(&v, g_trap_code, sizeof(v)) != cudaSuccess) return 0xFFFFFFFFu;
    if (v) cudaMemcpyToSymbol(g_trap_code, &z, sizeof(z));
    return v;
}

}  // namespace vsd
