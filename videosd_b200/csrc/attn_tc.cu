// Fused flash-style attention on tcgen05 tensor cores for the UNet's self- and cross-attention
// (replaces F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0).
//
// One CTA = 128 queries of one (image, head). Per 128-key block:
//   S = Q K^T          tcgen05.mma, fp32 accumulator in TMEM columns [0,128)
//   P = softmax piece  4 softmax warps read S with tcgen05.ld (thread = query row), online max / sum in fp32,
//                      write bf16 P into shared memory in the 128B-swizzled K-major operand layout
//   O += P V           tcgen05.mma, accumulator in TMEM columns [128, 128+dv), rescaled in place when the max moves
// Layouts (produced by the projection GEMMs, so no transposes happen here):
//   Q, K : [rows][heads * dk_pad] bf16, each head zero-padded from d to dk_pad (multiple of 64)
//   V^T  : [heads * d][keys]      bf16 (the V projection is computed with swapped operands)
// warp 0: TMA producer | warp 1: MMA issue + TMEM alloc | warps 2..9: softmax / correction / epilogue
// (row r of the tile is shared by warps w and w+4: TMEM lane quarter = warp % 4, each takes 64 of the 128 keys)
#include "tc_common.cuh"
#include "vsd_internal.h"
#include <algorithm>

namespace vsd {

static constexpr int kAttnThreads = 320;   // TMA warp, MMA warp, 8 softmax warps (two threads per query row)
static constexpr int kTileBytes = 128 * 128;  // one [128 rows][64 bf16] swizzled tile

struct AttnParams {
    int heads, d, dk_pad, dv_pad;
    int nq, nk;
    int q_rows_per_img, k_rows_per_img, vt_cols_per_img;
    bf16* out;
    int ldo;
    float scale_log2e;
    int stages, tmem_cols;
    int stagger;
    unsigned int* trace;
};

__device__ __forceinline__ float ex2_approx(float x) {   // one MUFU op; ex2(-inf) = 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapVt, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nkc = p.dk_pad >> 6;                       // 64-wide K chunks of the head dimension
    const uint32_t v_tile = (uint32_t)p.dv_pad * 128u;   // one [dv_pad][64 keys] tile
    const uint32_t stage_bytes = (uint32_t)nkc * kTileBytes + 2u * v_tile;
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + (size_t)nkc * kTileBytes;
    uint8_t* sP = sKV + (size_t)p.stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kTileBytes);
    uint64_t* q_full = bars;
    uint64_t* s_full = bars + 1;
    uint64_t* p_ready = bars + 2;
    uint64_t* pv_done = bars + 3;
    uint64_t* kv_full = bars + 4;
    uint64_t* kv_empty = kv_full + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_empty + p.stages);
    float* sx = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));  // [2][128] row exchange

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int nblocks = (p.nk + 127) >> 7;
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace, 1u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapK);
        tma_prefetch_desc(&mapVt);
        mbar_init(q_full, 1);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 256);
        mbar_init(pv_done, 1);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();               // only after this CTA owns its TMEM columns (see conv_gemm_kernel)
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 1, 1u);
    const uint32_t tS = tmem_base;         // 128 fp32 columns
    const uint32_t tO = tmem_base + 128u;  // dv_pad fp32 columns

    if (warp == 0) {
        // warp-uniform control flow; the elect.sync leader issues the TMA operations
        pdl_wait();   // Q / K / V^T are produced by the preceding projection kernels
        if (elect_one()) {
            mbar_expect_tx(q_full, (uint32_t)nkc * kTileBytes);
            for (int kc = 0; kc < nkc; ++kc)
                tma_load_2d(sQ + (size_t)kc * kTileBytes, &mapQ, q_full, h * p.dk_pad + kc * 64,
                            b * p.q_rows_per_img + qt * 128);
        }
        __syncwarp();
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < nblocks; ++j) {
            mbar_wait(&kv_empty[s], ph ^ 1u, 1);
            if (elect_one()) {
                mbar_expect_tx(&kv_full[s], stage_bytes);
                uint8_t* st = sKV + (size_t)s * stage_bytes;
                for (int kc = 0; kc < nkc; ++kc)
                    tma_load_2d(st + (size_t)kc * kTileBytes, &mapK, &kv_full[s], h * p.dk_pad + kc * 64,
                                b * p.k_rows_per_img + j * 128);
                uint8_t* sv = st + (size_t)nkc * kTileBytes;
                for (int a = 0; a < 2; ++a)
                    tma_load_2d(sv + (size_t)a * v_tile, &mapVt, &kv_full[s], b * p.vt_cols_per_img + j * 128 + a * 64,
                                h * p.d);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_bf16(128, 128);
        const uint32_t idesc_o = umma_idesc_bf16(128, (uint32_t)p.dv_pad);
        mbar_wait(q_full, 0, 2);
        int s = 0;
        uint32_t ph = 0;
        for (int j = 0; j < nblocks; ++j) {
            mbar_wait(&kv_full[s], ph, 3);
            tc_fence_after_sync();
            const uint32_t q_addr = smem_u32(sQ);
            const uint32_t k_addr = smem_u32(sKV + (size_t)s * stage_bytes);
            const uint32_t v_addr = k_addr + (uint32_t)nkc * kTileBytes;
            const uint32_t p_addr = smem_u32(sP);
            if (elect_one()) {
                const int nk16 = (p.d + 15) >> 4;   // columns past d are zero padding: skip their K-steps
                for (int k = 0; k < nk16; ++k) {
                    const uint32_t off = (uint32_t)(k >> 2) * kTileBytes + (uint32_t)(k & 3) * 32u;
                    umma_bf16(tS, umma_desc_sw128(q_addr + off), umma_desc_sw128(k_addr + off), idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(s_full);
            }
            __syncwarp();
            mbar_wait(p_ready, (uint32_t)j & 1u, 4);
            tc_fence_after_sync();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t aoff = (uint32_t)(k >> 2) * kTileBytes + (uint32_t)(k & 3) * 32u;
                    const uint32_t boff = (uint32_t)(k >> 2) * v_tile + (uint32_t)(k & 3) * 32u;
                    umma_bf16(tO, umma_desc_sw128(p_addr + aoff), umma_desc_sw128(v_addr + boff), idesc_o,
                              (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(&kv_empty[s]);
                umma_commit(pv_done);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else {
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;          // which 64 keys of the 128-key block this thread handles
        const int r = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        pdl_wait();   // the output buffer may still be read by an earlier kernel
        for (int j = 0; j < nblocks; ++j) {
            mbar_wait(s_full, (uint32_t)j & 1u, 5);
            tc_fence_after_sync();
            const int kbase = j * 128 + half * 64;
            const bool full_blk = (kbase + 64 <= p.nk);   // warp-uniform: no key masking needed
            // pass 1: row max over this thread's 64 keys, then combine with the partner thread of the row
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t u[32];
                tmem_ld32(tS + lane_off + half * 64 + c, u);
                tmem_ld_wait();
                if (full_blk) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) mx = fmaxf(mx, __uint_as_float(u[jj]));
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj)
                        if (kbase + c + jj < p.nk) mx = fmaxf(mx, __uint_as_float(u[jj]));
                }
            }
            sx[half * 128 + r] = mx;
            asm volatile("bar.sync 2, 256;" ::: "memory");
            mx = fmaxf(mx, sx[(half ^ 1) * 128 + r]);
            const float m_new = fmaxf(m_run, mx);
            const float alpha = ex2_approx((m_run - m_new) * p.scale_log2e);  // 0 on the first block
            if (j > 0) {
                // O and the P buffer belong to the previous P*V until it retires
                mbar_wait(pv_done, (uint32_t)(j - 1) & 1u, 6);
                tc_fence_after_sync();
                if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
                    for (int c = half * 16; c < p.dv_pad; c += 32) {   // the two threads of a row interleave 16-col chunks
                        uint32_t o[16];
                        tmem_ld16(tO + lane_off + c, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
                        tmem_st16(tO + lane_off + c, o);
                    }
                    tmem_st_wait();
                }
            }
            l_run *= alpha;
            // pass 2: probabilities -> bf16, one 128-byte row of P tile atom `half` (swizzled K-major layout)
            float lsum = 0.f;
            const float mscaled = m_new * p.scale_log2e;
            const uint32_t prow = smem_u32(sP) + (uint32_t)half * kTileBytes + (uint32_t)r * 128u;
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t u[32];
                tmem_ld32(tS + lane_off + half * 64 + c, u);
                tmem_ld_wait();
                float pv[32];
                if (full_blk) {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const float e = ex2_approx(fmaf(__uint_as_float(u[jj]), p.scale_log2e, -mscaled));
                        lsum += e;
                        pv[jj] = e;
                    }
                } else {
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        float e = ex2_approx(fmaf(__uint_as_float(u[jj]), p.scale_log2e, -mscaled));
                        e = (kbase + c + jj < p.nk) ? e : 0.f;
                        lsum += e;
                        pv[jj] = e;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int chunk = (c >> 3) + i;  // 16-byte chunk inside the 128-byte row
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (uint32_t)((chunk ^ (r & 7)) << 4)),
                                 "r"(pack_bf16x2(pv[i * 8], pv[i * 8 + 1])), "r"(pack_bf16x2(pv[i * 8 + 2], pv[i * 8 + 3])),
                                 "r"(pack_bf16x2(pv[i * 8 + 4], pv[i * 8 + 5])), "r"(pack_bf16x2(pv[i * 8 + 6], pv[i * 8 + 7])) : "memory");
                }
            }
            l_run += lsum;
            m_run = m_new;
            tc_fence_before_sync();
            fence_proxy_async_smem();
            mbar_arrive(p_ready);
        }
        mbar_wait(pv_done, (uint32_t)(nblocks - 1) & 1u, 7);
        tc_fence_after_sync();
        sx[half * 128 + r] = l_run;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        const float inv_l = 1.0f / (l_run + sx[(half ^ 1) * 128 + r]);
        const int qrow = qt * 128 + r;
        const bool row_ok = qrow < p.nq;
        bf16* orow = p.out + ((long)b * p.nq + qrow) * p.ldo + h * p.d;
#pragma unroll 1
        for (int c = half * 16; c < p.d; c += 32) {
            uint32_t o[16];
            tmem_ld16(tO + lane_off + c, o);
            tmem_ld_wait();
            float f[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) f[jj] = __uint_as_float(o[jj]) * inv_l;
            if (row_ok) {
                *reinterpret_cast<uint4*>(orow + c) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                                 pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                if (c + 8 < p.d)
                    *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                                                                         pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 2, 1u);
}

// ------------------------------------------------------------------------------------------ two-tile kernel
// attention2_kernel: the kernel for the long self-attention layers (head dim <= 64, >= 2 key blocks, >= 256 queries).
// One CTA = 256 queries of one (image, head) as TWO 128-row tiles with their own S / O accumulators in TMEM
// (S0 | S1 | O0 | O1 = 512 columns) and their own softmax warpgroup, so that the tensor core computes one tile's
// Q K^T / P V while the other tile's softmax runs; per tile the next block's Q K^T is issued as soon as the softmax
// threads hold the current S row in registers (s_free), i.e. it overlaps the exponentials as well.
//   * one thread per query row: the whole 128-key row of S is read from TMEM ONCE, no cross-thread row exchange;
//   * P is DOUBLE-BUFFERED per tile: the softmax of block j writes P[j & 1] while P V of block j - 1 still reads the other
//     buffer, so a warpgroup never waits for the tensor core between two blocks (ncu of the single-buffer version: 19 % of the
//     softmax warps' time was the wait for the previous P V; profiles/r02_attention2.md);
//   * lazy rescale: the running maximum only moves when the block maximum exceeds it by more than 2^8 (P <= 256 is
//     harmless in bf16 / fp32), so the TMEM read-modify-write of O (which does need the previous P V) happens in the first
//     blocks only;
//   * K and V^T have their own rings: a K tile is dead as soon as its Q K^T retires (a block earlier than the V^T tile), so the
//     next K arrives a whole block ahead with only two slots each;
//   * optional: every kPoly-th exponential on the FMA pipe (Cody-Waite reduction + degree-3 polynomial, relative error 8e-5)
//     -- measured slower than all-MUFU at this head dimension (the kernel is not MUFU-bound enough), kept as a knob;
//   * Q K^T runs over ceil(d / 16) K-steps (48 columns for d = 40), not the 64-column padding.
// warp 0: TMA | warps 1, 2: MMA issue for tile 0 / tile 1 (warp 1 also allocates TMEM) | warp 3: idle |
// warps 4..7: softmax of tile 0 | warps 8..11: softmax of tile 1
static constexpr int kAttn2Threads = 384;   // TMA warp, one MMA warp per tile, an idle warp, two softmax warpgroups

__device__ __forceinline__ float max3f(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.f);
    const float r = x + 12582912.f;            // 1.5 * 2^23: round(x) lands in the low mantissa bits
    const float f = x - (r - 12582912.f);      // [-0.5, 0.5]
    float q = fmaf(0.05508905f, f, 0.24260408f);
    q = fmaf(q, f, 0.69327617f);
    q = fmaf(q, f, 0.99992895f);
    return __int_as_float(__float_as_int(q) + (__float_as_int(r) << 23));
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// exponentials of one 32-key chunk of a row -> bf16 -> the row's 64 bytes in the swizzled P tile; returns their sum
// prow_atom: shared-space address of this row inside the [128][64] atom
template <int kPoly>
__device__ __forceinline__ float softmax_chunk32(const uint32_t (&u)[32], float sc, float ms, uint32_t prow_atom, int chunk0, int rsw) {
    float e[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
        const float x = fmaf(__uint_as_float(u[jj]), sc, -ms);
        e[jj] = (kPoly > 0 && (jj % (kPoly > 0 ? kPoly : 1)) == (kPoly - 1)) ? ex2_poly(x) : ex2_approx(x);
    }
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int jj = 0; jj < 32; jj += 4) { s0 += e[jj]; s1 += e[jj + 1]; s2 += e[jj + 2]; s3 += e[jj + 3]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        sts128(prow_atom + (uint32_t)(((chunk0 + i) ^ rsw) << 4), pack_bf16x2(e[i * 8], e[i * 8 + 1]), pack_bf16x2(e[i * 8 + 2], e[i * 8 + 3]),
               pack_bf16x2(e[i * 8 + 4], e[i * 8 + 5]), pack_bf16x2(e[i * 8 + 6], e[i * 8 + 7]));
    return (s0 + s1) + (s2 + s3);
}

__device__ __forceinline__ float rowmax32(const uint32_t (&u)[32], float m) {
    float a = m, b = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 32; jj += 4) {
        a = max3f(a, __uint_as_float(u[jj]), __uint_as_float(u[jj + 1]));
        b = max3f(b, __uint_as_float(u[jj + 2]), __uint_as_float(u[jj + 3]));
    }
    return fmaxf(a, b);
}

__device__ __forceinline__ void mask32(uint32_t (&u)[32], int valid) {   // keys >= valid -> -inf
#pragma unroll
    for (int jj = 0; jj < 32; ++jj)
        if (jj >= valid) u[jj] = 0xff800000u;
}

template <int kPoly>
__global__ void __launch_bounds__(kAttn2Threads, 1)
attention2_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                  const __grid_constant__ CUtensorMap mapVt, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int KS = 2, VS = 2;                               // K / V^T ring depths
    const uint32_t v_tile = (uint32_t)p.dv_pad * 128u;          // one [dv_pad][64 keys] tile
    uint8_t* sQ = smem;                                         // [2 tiles][128][64]
    uint8_t* sP = sQ + 2 * kTileBytes;                          // [2 tiles][2 buffers][2 atoms][128][64]
    uint8_t* sK = sP + 8 * kTileBytes;                          // [KS][128 keys][64]
    uint8_t* sV = sK + KS * kTileBytes;                         // [VS][2][dv_pad][64 keys]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + (size_t)VS * 2 * v_tile);
    uint64_t* q_full = bars;
    uint64_t* s_full = bars + 1;     // [tile]
    uint64_t* s_free = bars + 3;     // [tile]
    uint64_t* p_ready = bars + 5;    // [tile][buffer]
    uint64_t* pv_done = bars + 9;    // [tile][buffer]
    uint64_t* k_full = bars + 13;    // [KS]
    uint64_t* k_empty = bars + 15;
    uint64_t* v_full = bars + 17;    // [VS]
    uint64_t* v_empty = bars + 19;
    uint64_t* stagger = bars + 21;   // tile 0's warpgroup has reached its first exponentials
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int nblocks = (p.nk + 127) >> 7;
    const int ntiles = (qt * 256 + 128 < p.nq) ? 2 : 1;         // the second tile of the last CTA may lie past the queries
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace, 1u);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapK);
        tma_prefetch_desc(&mapVt);
        mbar_init(q_full, 1);
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 4);
            for (int k = 0; k < 2; ++k) {
                mbar_init(&p_ready[t * 2 + k], 4);
                mbar_init(&pv_done[t * 2 + k], 1);
            }
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], (uint32_t)ntiles);           // one tcgen05.commit per tile's MMA warp
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], (uint32_t)ntiles);
        }
        mbar_init(stagger, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512u);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();   // only after this CTA owns its TMEM columns (see conv_gemm_kernel)
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 1, 1u);

    if (warp == 0) {
        pdl_wait();   // Q / K / V^T are produced by the preceding projection kernels
        if (elect_one()) {
            mbar_expect_tx(q_full, (uint32_t)ntiles * kTileBytes);
            for (int t = 0; t < ntiles; ++t)
                tma_load_2d(sQ + (size_t)t * kTileBytes, &mapQ, q_full, h * p.dk_pad, b * p.q_rows_per_img + qt * 256 + t * 128);
            auto load_k = [&](int j) {
                const int s = j % KS;
                if (j >= KS) mbar_wait(&k_empty[s], (uint32_t)((j / KS) - 1) & 1u, 1);
                mbar_expect_tx(&k_full[s], (uint32_t)kTileBytes);
                tma_load_2d(sK + (size_t)s * kTileBytes, &mapK, &k_full[s], h * p.dk_pad, b * p.k_rows_per_img + j * 128);
            };
            auto load_v = [&](int j) {
                const int s = j % VS;
                if (j >= VS) mbar_wait(&v_empty[s], (uint32_t)((j / VS) - 1) & 1u, 2);
                mbar_expect_tx(&v_full[s], 2u * v_tile);
                for (int a = 0; a < 2; ++a)
                    tma_load_2d(sV + ((size_t)s * 2 + a) * v_tile, &mapVt, &v_full[s], b * p.vt_cols_per_img + j * 128 + a * 64, h * p.d);
            };
            // K runs one block ahead of V^T (it is needed a block earlier): K0, K1, V0, K2, V1, ...
            load_k(0);
            for (int j = 0; j < nblocks; ++j) {
                if (j + 1 < nblocks) load_k(j + 1);
                load_v(j);
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2) {
        // One MMA warp PER TILE (ncu of the one-warp version: that warp was busy 85 % of the time -- 22 tcgen05.mma, 6 commits
        // and 6 barrier polls per key block in one instruction stream -- and the softmax warps waited for it). A single elected
        // thread runs the whole loop: waits, descriptor arithmetic in uniform registers, issue.
        const int t = warp - 1;
        if (t < ntiles && elect_one()) {
            const uint32_t idesc_s = umma_idesc_bf16(128, 128);
            const uint32_t idesc_o = umma_idesc_bf16(128, (uint32_t)p.dv_pad);
            const int nk16 = (p.d + 15) >> 4;                   // K-steps of Q K^T (columns past d are zero padding)
            // shared-memory descriptors differ only in their 14-bit start-address field ((addr & 0x3FFFF) >> 4)
            const uint64_t dq = umma_desc_sw128(smem_u32(sQ) + (uint32_t)t * kTileBytes), dk = umma_desc_sw128(smem_u32(sK)),
                           dp = umma_desc_sw128(smem_u32(sP) + (uint32_t)t * 4u * kTileBytes), dv = umma_desc_sw128(smem_u32(sV));
            const uint32_t tS = tmem_base + (uint32_t)t * 128u, tO = tmem_base + 256u + (uint32_t)t * 128u;
            auto issue_qk = [&](int j) {                        // S_t = Q_t K(j)^T
                const uint64_t b0 = dk + (uint64_t)((uint32_t)(j % KS) * (kTileBytes >> 4));
                for (int k = 0; k < nk16; ++k) umma_bf16(tS, dq + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
                umma_commit(&s_full[t]);
                umma_commit(&k_empty[j % KS]);
            };
            mbar_wait(q_full, 0, 3);
            mbar_wait(&k_full[0], 0, 4);
            tc_fence_after_sync();
            issue_qk(0);
            for (int j = 0; j < nblocks; ++j) {
                if (j + 1 < nblocks) {
                    mbar_wait(&k_full[(j + 1) % KS], (uint32_t)((j + 1) / KS) & 1u, 4);
                    mbar_wait(&s_free[t], (uint32_t)j & 1u, 5);   // the softmax threads hold S_t(j) in registers
                    tc_fence_after_sync();
                    issue_qk(j + 1);
                }
                mbar_wait(&v_full[j % VS], (uint32_t)(j / VS) & 1u, 6);
                mbar_wait(&p_ready[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1u, 7);
                tc_fence_after_sync();
                const uint64_t a0 = dp + (uint64_t)((uint32_t)(j & 1) * ((2u * kTileBytes) >> 4));
                const uint64_t b0 = dv + (uint64_t)((uint32_t)(j % VS) * ((2u * v_tile) >> 4));
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tO, a0 + (uint64_t)((k >> 2) * (kTileBytes >> 4) + (k & 3) * 2),
                              b0 + (uint64_t)((uint32_t)(k >> 2) * (v_tile >> 4) + (uint32_t)(k & 3) * 2u), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                umma_commit(&pv_done[t * 2 + (j & 1)]);
                umma_commit(&v_empty[j % VS]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int t = (warp - 4) >> 2;                          // tile of this warpgroup
        const int q = warp & 3;                                 // TMEM lane quarter
        const int r = q * 32 + lane;
        if (t < ntiles) {
            const uint32_t lane_off = (uint32_t)(q * 32) << 16;
            const uint32_t tS = tmem_base + lane_off + (uint32_t)t * 128u;
            const uint32_t tO = tmem_base + lane_off + 256u + (uint32_t)t * 128u;
            const uint32_t prow_t = smem_u32(sP) + (uint32_t)t * 4u * kTileBytes + (uint32_t)r * 128u;
            const int rsw = r & 7;
            const float sc = p.scale_log2e;
            float m_used = -INFINITY, l_run = 0.f;
            pdl_wait();   // the output buffer may still be read by an earlier kernel
            // The two warpgroups share each scheduler's MUFU pipe. Started together they load S / search the row maximum at the
            // same time (MUFU idle) and then contend for the exponentials; tile 1 therefore starts when tile 0 reaches its first
            // exponentials, so that one group's S load + maximum overlaps the other's exponentials (VSD_ATTN_STAGGER knob).
            if (t == 1 && p.stagger) mbar_wait(stagger, 0, 12);
            for (int j = 0; j < nblocks; ++j) {
                mbar_wait(&s_full[t], (uint32_t)j & 1u, 8);
                tc_fence_after_sync();
                uint32_t u0[32], u1[32], u2[32], u3[32];
                tmem_ld32(tS, u0);
                tmem_ld32(tS + 32, u1);
                tmem_ld32(tS + 64, u2);
                tmem_ld32(tS + 96, u3);
                tmem_ld_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[t]);         // the next Q K^T of this tile may overwrite S
                const int valid = p.nk - j * 128;               // warp-uniform
                if (valid < 128) { mask32(u0, valid); mask32(u1, valid - 32); mask32(u2, valid - 64); mask32(u3, valid - 96); }
                const float mx = rowmax32(u3, rowmax32(u2, rowmax32(u1, rowmax32(u0, -INFINITY))));
                if (j == 0) {
                    m_used = mx;
                    if (t == 0 && lane == 0) mbar_arrive(stagger);
                } else {
                    const bool grow = (mx - m_used) * sc > 8.0f;
                    if (__any_sync(0xffffffffu, grow)) {
                        // O_t belongs to the previous P V until it retires
                        mbar_wait(&pv_done[t * 2 + ((j - 1) & 1)], (uint32_t)((j - 1) >> 1) & 1u, 9);
                        tc_fence_after_sync();
                        const float m_new = grow ? mx : m_used;
                        const float alpha = ex2_approx((m_used - m_new) * sc);   // 1 for the rows that keep their maximum
                        m_used = m_new;
                        l_run *= alpha;
#pragma unroll 1
                        for (int c = 0; c < p.dv_pad; c += 16) {
                            uint32_t o[16];
                            tmem_ld16(tO + c, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
                            tmem_st16(tO + c, o);
                        }
                        tmem_st_wait();
                    }
                    // this block's P buffer was last read by the P V of block j - 2
                    if (j >= 2) mbar_wait(&pv_done[t * 2 + (j & 1)], (uint32_t)((j >> 1) - 1) & 1u, 10);
                }
                const float ms = m_used * sc;
                const uint32_t prow = prow_t + (uint32_t)(j & 1) * 2u * kTileBytes;
                float lsum = softmax_chunk32<kPoly>(u0, sc, ms, prow, 0, rsw);
                lsum += softmax_chunk32<kPoly>(u1, sc, ms, prow, 4, rsw);
                lsum += softmax_chunk32<kPoly>(u2, sc, ms, prow + kTileBytes, 0, rsw);
                lsum += softmax_chunk32<kPoly>(u3, sc, ms, prow + kTileBytes, 4, rsw);
                l_run += lsum;
                tc_fence_before_sync();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_ready[t * 2 + (j & 1)]);
            }
            mbar_wait(&pv_done[t * 2 + ((nblocks - 1) & 1)], (uint32_t)((nblocks - 1) >> 1) & 1u, 11);
            tc_fence_after_sync();
            const float inv_l = 1.0f / l_run;
            const int qrow = qt * 256 + t * 128 + r;
            const bool row_ok = qrow < p.nq;
            bf16* orow = p.out + ((long)b * p.nq + qrow) * p.ldo + h * p.d;
#pragma unroll 1
            for (int c = 0; c < p.d; c += 16) {
                uint32_t o[16];
                tmem_ld16(tO + c, o);
                tmem_ld_wait();
                float f[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) f[jj] = __uint_as_float(o[jj]) * inv_l;
                if (row_ok) {
                    *reinterpret_cast<uint4*>(orow + c) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                                     pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    if (c + 8 < p.d)
                        *reinterpret_cast<uint4*>(orow + c + 8) = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                                                                             pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512u);
    if (p.trace && threadIdx.x == 0) atomicAdd(p.trace + 2, 1u);
}

// ------------------------------------------------------------------------------------------ host side
static int g_attn_max_smem = 227 * 1024;

int attn_init() {
    int dev = 0;
    VSD_CHECK_CUDA(cudaGetDevice(&dev));
    VSD_CHECK_CUDA(cudaDeviceGetAttribute(&g_attn_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    VSD_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_attn_max_smem));
    VSD_CHECK_CUDA(cudaFuncSetAttribute(attention2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_attn_max_smem));
    VSD_CHECK_CUDA(cudaFuncSetAttribute(attention2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_attn_max_smem));
    VSD_CHECK_CUDA(cudaFuncSetAttribute(attention2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_attn_max_smem));
    VSD_CHECK_CUDA(cudaFuncSetAttribute(attention2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_attn_max_smem));
    return 0;
}

int attn_dk_pad(int d) { return ((d + 63) / 64) * 64; }
int attn_dv_pad(int d) { return ((d + 15) / 16) * 16; }

int build_attn_op(AttnOp* op, const bf16* q, int ldq, const bf16* k, int ldk, const bf16* vt, int ldvt, bf16* out,
                  int ldo, int batch, int heads, int d, int nq, int nk, int q_rows_per_img, int k_rows_per_img,
                  int vt_cols_per_img, int vt_rows) {
    VSD_REQUIRE(d % 8 == 0 && d <= 256, "head dim must be a multiple of 8 and <= 256");
    VSD_REQUIRE(nq > 0 && nk > 0, "empty attention");
    // the V^T tiles start at column b * vt_cols_per_img + j * 128: TMA needs 16-byte aligned box origins along the contiguous dimension
    VSD_REQUIRE(batch == 1 || (vt_cols_per_img % 8) == 0, "V^T columns per image must be a multiple of 8 when batch > 1");
    op->heads = heads; op->d = d; op->dk_pad = attn_dk_pad(d); op->dv_pad = attn_dv_pad(d);
    op->nq = nq; op->nk = nk; op->batch = batch;
    op->q_rows_per_img = q_rows_per_img; op->k_rows_per_img = k_rows_per_img; op->vt_cols_per_img = vt_cols_per_img;
    op->out = out; op->ldo = ldo;
    op->scale_log2e = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
    VSD_REQUIRE((ldo % 8) == 0 && ((heads * d) <= ldo), "attention output stride");
    int rc = make_tmap_2d(&op->mapQ, q, heads * op->dk_pad, batch * q_rows_per_img, ldq, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&op->mapK, k, heads * op->dk_pad, batch * k_rows_per_img, ldk, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&op->mapVt, vt, batch * vt_cols_per_img, vt_rows, ldvt, op->dv_pad);
    if (rc) return rc;
    // Long self-attention (head dim <= 64): the two-tile kernel (attention2_kernel). VSD_ATTN_V2=0 keeps the one-tile kernel,
    // VSD_ATTN_POLY = n: every n-th exponential on the FMA pipe (0: all on the MUFU pipe).
    static const bool v2_ok = !(getenv("VSD_ATTN_V2") && atoi(getenv("VSD_ATTN_V2")) == 0);
    static const int poly = getenv("VSD_ATTN_POLY") ? atoi(getenv("VSD_ATTN_POLY")) : 0;
    op->variant = 0;
    op->trace = nullptr;
    if (v2_ok && op->dk_pad == 64 && nk > 128 && nq > 128) {
        // Q (2 tiles) + P (2 tiles x 2 buffers x 2 atoms) + K ring (2) + V^T ring (2 x 2 x dv_pad x 128 B) + barriers
        const int need2 = 12 * kTileBytes + 4 * op->dv_pad * 128 + 1024 /*align slack*/ + 512 /*barriers*/;
        if (need2 <= g_attn_max_smem) {
            op->variant = 2;
            op->poly = (poly == 2 || poly == 3 || poly == 4) ? poly : 0;
            op->stages = 2;
            op->tmem_cols = 512;
            op->smem_bytes = std::max(need2, std::min(512 * 450, g_attn_max_smem));   // 512 TMEM columns: the SM is ours
            op->grid = dim3((nq + 255) / 256, heads, batch);
            return 0;
        }
    }
    const int nkc = op->dk_pad / 64;
    const int stage_bytes = nkc * kTileBytes + 2 * op->dv_pad * 128;
    const int fixed = nkc * kTileBytes + 2 * kTileBytes + 1024 + 256 + 1024 + 32;   // + row-exchange scratch
    int stages = 2;
    if (fixed + stages * stage_bytes > g_attn_max_smem) stages = 1;
    VSD_REQUIRE(fixed + stages * stage_bytes <= g_attn_max_smem, "attention tile does not fit shared memory");
    const int nblocks = (nk + 127) / 128;
    if (stages > nblocks) stages = nblocks;
    op->stages = stages;
    op->smem_bytes = fixed + stages * stage_bytes;
    const int cols = 128 + op->dv_pad;
    op->tmem_cols = cols <= 256 ? 256 : 512;
    // TMEM over-subscription guard under programmatic dependent launch (see build_gemm_op)
    if (op->smem_bytes < op->tmem_cols * 450) op->smem_bytes = std::min(op->tmem_cols * 450, g_attn_max_smem);
    op->grid = dim3((nq + 127) / 128, heads, batch);
    return 0;
}

int launch_attn_op(const AttnOp& op, cudaStream_t st) {
    AttnParams p;
    p.heads = op.heads; p.d = op.d; p.dk_pad = op.dk_pad; p.dv_pad = op.dv_pad;
    p.nq = op.nq; p.nk = op.nk;
    p.q_rows_per_img = op.q_rows_per_img; p.k_rows_per_img = op.k_rows_per_img; p.vt_cols_per_img = op.vt_cols_per_img;
    p.out = op.out; p.ldo = op.ldo; p.scale_log2e = op.scale_log2e; p.stages = op.stages; p.tmem_cols = op.tmem_cols;
    static const int stagger = getenv("VSD_ATTN_STAGGER") ? atoi(getenv("VSD_ATTN_STAGGER")) : 0;   // measured: no effect either way
    p.stagger = stagger;
    p.trace = op.trace;
    if (op.variant == 2) {
        auto kern = op.poly == 2 ? attention2_kernel<2> : (op.poly == 3 ? attention2_kernel<3> : (op.poly == 4 ? attention2_kernel<4> : attention2_kernel<0>));
        VSD_CHECK_CUDA(launch_k(kern, op.grid, dim3(kAttn2Threads), (size_t)op.smem_bytes, st, op.mapQ, op.mapK, op.mapVt, p));
        return 0;
    }
    VSD_CHECK_CUDA(launch_k(attention_kernel, op.grid, dim3(kAttnThreads), (size_t)op.smem_bytes, st, op.mapQ, op.mapK, op.mapVt, p));
    return 0;
}

const unsigned int* trap_code_addr_attn() {
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_trap_code) == cudaSuccess ? static_cast<const unsigned int*>(p) : nullptr;
}

unsigned int read_trap_code_attn() {
    unsigned int v = 0, z = 0;
    if (cudaMemcpyFromSymbol(&v, g_trap_code, sizeof(v)) != cudaSuccess) return 0xFFFFFFFFu;
    if (v) cudaMemcpyToSymbol(g_trap_code, &z, sizeof(z));
    return v;
}

}  // namespace vsd
