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
    pdl_launch_dependents();
    const int nblocks = (p.nk + 127) >> 7;

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
                const int nk16 = p.dk_pad >> 4;
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
            uint8_t* prow = sP + (size_t)half * kTileBytes + (size_t)r * 128;
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
                    const uint4 w = make_uint4(pack_bf16x2(pv[i * 8], pv[i * 8 + 1]), pack_bf16x2(pv[i * 8 + 2], pv[i * 8 + 3]),
                                               pack_bf16x2(pv[i * 8 + 4], pv[i * 8 + 5]), pack_bf16x2(pv[i * 8 + 6], pv[i * 8 + 7]));
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ (r & 7)) << 4)) = w;
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
}

// ------------------------------------------------------------------------------------------ host side
static int g_attn_max_smem = 227 * 1024;

int attn_init() {
    int dev = 0;
    VSD_CHECK_CUDA(cudaGetDevice(&dev));
    VSD_CHECK_CUDA(cudaDeviceGetAttribute(&g_attn_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    VSD_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g_attn_max_smem));
    return 0;
}

int attn_dk_pad(int d) { return ((d + 63) / 64) * 64; }
int attn_dv_pad(int d) { return ((d + 15) / 16) * 16; }

int build_attn_op(AttnOp* op, const bf16* q, int ldq, const bf16* k, int ldk, const bf16* vt, int ldvt, bf16* out,
                  int ldo, int batch, int heads, int d, int nq, int nk, int q_rows_per_img, int k_rows_per_img,
                  int vt_cols_per_img, int vt_rows) {
    VSD_REQUIRE(d % 8 == 0 && d <= 256, "head dim must be a multiple of 8 and <= 256");
    VSD_REQUIRE(nq > 0 && nk > 0, "empty attention");
    op->heads = heads; op->d = d; op->dk_pad = attn_dk_pad(d); op->dv_pad = attn_dv_pad(d);
    op->nq = nq; op->nk = nk; op->batch = batch;
    op->q_rows_per_img = q_rows_per_img; op->k_rows_per_img = k_rows_per_img; op->vt_cols_per_img = vt_cols_per_img;
    op->out = out; op->ldo = ldo;
    op->scale_log2e = (1.0f / sqrtf((float)d)) * 1.4426950408889634f;
    VSD_REQUIRE((ldo % 8) == 0 && ((heads * d) <= ldo), "attention output stride");
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
    int rc = make_tmap_2d(&op->mapQ, q, heads * op->dk_pad, batch * q_rows_per_img, ldq, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&op->mapK, k, heads * op->dk_pad, batch * k_rows_per_img, ldk, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&op->mapVt, vt, batch * vt_cols_per_img, vt_rows, ldvt, op->dv_pad);
    if (rc) return rc;
    op->grid = dim3((nq + 127) / 128, heads, batch);
    return 0;
}

int launch_attn_op(const AttnOp& op, cudaStream_t st) {
    AttnParams p;
    p.heads = op.heads; p.d = op.d; p.dk_pad = op.dk_pad; p.dv_pad = op.dv_pad;
    p.nq = op.nq; p.nk = op.nk;
    p.q_rows_per_img = op.q_rows_per_img; p.k_rows_per_img = op.k_rows_per_img; p.vt_cols_per_img = op.vt_cols_per_img;
    p.out = op.out; p.ldo = op.ldo; p.scale_log2e = op.scale_log2e; p.stages = op.stages; p.tmem_cols = op.tmem_cols;
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
