// Encoder self-attention (non-causal, d_head = 64) as a flash-style kernel on tcgen05 tensor cores.
// Replaces CTranslate2's batched-cuBLAS + softmax attention inside ctranslate2.models.Whisper.encode
// (SURVEY.md §8 a6, A.7).
//
// One CTA = one (batch, head, 128-query tile); 128 threads, thread i owns query row i.
//   S = Q K_j^T        tcgen05.mma 128x128x64  (Q, K_j: TMA, K-major SWIZZLE_128B)       -> TMEM cols [0,128)
//   softmax            thread reads its S row with tcgen05.ld (two passes: max, then exp2), writes the
//                      un-normalised P as bf16 into 128B-swizzled shared memory (A operand of the next MMA)
//   O_j = P V_j        tcgen05.mma 128x64x128  (V_j: TMA tile [keys, d] used MN-major)     -> TMEM cols [128,192)
//   acc = acc*alpha + O_j in registers (fp32), final acc / l -> bf16.
// Two CTAs fit per SM (96 KB smem, 256 TMEM columns each), so one CTA's softmax overlaps the other's MMAs.
#include "gemm.cuh"
#include "ptx_sm100.cuh"

namespace mw {
using namespace ptx;

namespace {

constexpr int TQ = 128;   // queries per CTA
constexpr int TK = 128;   // keys per block
constexpr int DH = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;   // 16 KB
constexpr int ATT_SMEM = 6 * TILE_BYTES + 128 + 1024;   // Q, K0, K1, V, P(2 tiles) + barriers + align slack

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(128, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ out,
                         int T, int d_model, float scale_log2e) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sQ = smem;
    unsigned char* sK = smem + TILE_BYTES;            // two stages
    unsigned char* sV = smem + 3 * TILE_BYTES;
    unsigned char* sP = smem + 4 * TILE_BYTES;        // two 64-key halves
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
    uint64_t* bar_q = bars;
    uint64_t* bar_k = bars + 1;   // [2]
    uint64_t* bar_v = bars + 3;
    uint64_t* bar_s = bars + 4;
    uint64_t* bar_o = bars + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int q0 = blockIdx.x * TQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int n_blocks = (T + TK - 1) / TK;
    const int col_q = h * DH, col_k = d_model + h * DH, col_v = 2 * d_model + h * DH;

    if (tid == 0) {
        prefetch_tensormap(&tmap_qkv);
        for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + 128;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

    if (tid == 0) {
        mbar_arrive_expect_tx(bar_q, TILE_BYTES);
        tma_load_3d(sQ, &tmap_qkv, bar_q, col_q, q0, b);
        mbar_arrive_expect_tx(&bar_k[0], TILE_BYTES);
        tma_load_3d(sK, &tmap_qkv, &bar_k[0], col_k, 0, b);
        if (n_blocks > 1) {
            mbar_arrive_expect_tx(&bar_k[1], TILE_BYTES);
            tma_load_3d(sK + TILE_BYTES, &tmap_qkv, &bar_k[1], col_k, TK, b);
        }
        mbar_arrive_expect_tx(bar_v, TILE_BYTES);
        tma_load_3d(sV, &tmap_qkv, bar_v, col_v, 0, b);
    }

    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 1);   // B (=V) is MN-major

    float acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) acc[i] = 0.0f;
    float m_run = -INFINITY, l_run = 0.0f;

    for (int j = 0; j < n_blocks; ++j) {
        const int ks = j & 1;
        if (tid == 0) {
            if (j == 0) mbar_wait(bar_q, 0);
            mbar_wait(&bar_k[ks], (j >> 1) & 1);
            tc_fence_after();
            const uint64_t dq = make_desc_sw128(smem_u32(sQ), 1024, 0);
            const uint64_t dk = make_desc_sw128(smem_u32(sK + ks * TILE_BYTES), 1024, 0);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
            umma_commit(bar_s);
        }
        __syncwarp();
        mbar_wait(bar_s, j & 1);
        tc_fence_after();
        if (tid == 0 && j + 2 < n_blocks) {      // K stage `ks` is free again
            mbar_arrive_expect_tx(&bar_k[ks], TILE_BYTES);
            tma_load_3d(sK + ks * TILE_BYTES, &tmap_qkv, &bar_k[ks], col_k, (j + 2) * TK, b);
        }
        __syncwarp();
        const int key0 = j * TK;
        const int n_valid = min(TK, T - key0);    // >= 1
        // ---- pass 1: row max
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < TK; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_s + lane_off + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c + i < n_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
        const float m_new = fmaxf(m_run, mx);
        const float alpha = ex2((m_run - m_new) * scale_log2e);   // m_run = -inf on the first block -> 0
        const float mb = m_new * scale_log2e;
        // ---- pass 2: p = exp2(s*c - m*c), row sum, bf16 P into swizzled smem
        float lsum = 0.0f;
#pragma unroll 1
        for (int c = 0; c < TK; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_s + lane_off + c, r);
            tmem_ld_wait();
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float p0 = (c + i < n_valid) ? ex2(fmaf(__uint_as_float(r[i]), scale_log2e, -mb)) : 0.0f;
                float p1 = (c + i + 1 < n_valid) ? ex2(fmaf(__uint_as_float(r[i + 1]), scale_log2e, -mb)) : 0.0f;
                __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
                // the row sum uses the rounded values the tensor core will see
                lsum += __low2float(hh) + __high2float(hh);
                packed[i >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
            unsigned char* prow = sP + (c >> 6) * TILE_BYTES + tid * 128;
            const int cbase = ((c & 63) >> 3);       // first 16-byte chunk of this 32-key group inside the 128-byte row
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int chunk = (cbase + g) ^ (tid & 7);
                *reinterpret_cast<uint4*>(prow + chunk * 16) =
                    make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
            }
        }
        l_run = l_run * alpha + lsum;
        m_run = m_new;
        fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core's async proxy
        tc_fence_before();
        __syncthreads();          // P complete, every S read retired
        if (tid == 0) {
            mbar_wait(bar_v, j & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TK / 16; ++k) {
                const uint64_t dp = make_desc_sw128(smem_u32(sP + (k >> 2) * TILE_BYTES) + (k & 3) * 32, 1024, 0);
                const uint64_t dv = make_desc_sw128(smem_u32(sV) + k * 2048, 1024, TILE_BYTES);
                umma_bf16(tmem_o, dp, dv, idesc_o, k ? 1u : 0u);
            }
            umma_commit(bar_o);
        }
        __syncwarp();
        mbar_wait(bar_o, j & 1);
        tc_fence_after();
        if (tid == 0 && j + 1 < n_blocks) {      // V buffer is free again
            mbar_arrive_expect_tx(bar_v, TILE_BYTES);
            tma_load_3d(sV, &tmap_qkv, bar_v, col_v, (j + 1) * TK, b);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < DH; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_o + lane_off + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[c + i] = fmaf(acc[c + i], alpha, __uint_as_float(r[i]));
        }
        tc_fence_before();
    }

    const int q = q0 + tid;
    if (q < T) {
        const float inv = 1.0f / l_run;
        uint4* o4 = reinterpret_cast<uint4*>(out + ((int64_t)b * T + q) * d_model + h * DH);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 hh = __floats2bfloat162_rn(acc[8 * g + 2 * i] * inv, acc[8 * g + 2 * i + 1] * inv);
                w[i] = *reinterpret_cast<uint32_t*>(&hh);
            }
            o4[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

mw_status attention_launch(const void* d_qkv, void* d_out, int B, int T, int n_heads, cudaStream_t st) {
    MW_REQUIRE(d_qkv && d_out && B > 0 && T > 0 && n_heads > 0, "attention: bad arguments");
    const int d = n_heads * DH;
    CUtensorMap tm;
    uint64_t dims[3] = {(uint64_t)3 * d, (uint64_t)T, (uint64_t)B};
    uint64_t str[2] = {(uint64_t)3 * d * 2, (uint64_t)T * 3 * d * 2};
    uint32_t box[3] = {DH, 128, 1};
    mw_status s = encode_tensor_map(&tm, d_qkv, 3, dims, str, box, true);
    if (s != MW_OK) return s;
    static bool attr_set = false;
    if (!attr_set) {
        MW_CUDA_CHECK(cudaFuncSetAttribute(attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set = true;
    }
    dim3 grid(ceil_div(T, TQ), n_heads, B);
    const float scale_log2e = 0.125f * 1.4426950408889634f;   // d_head^-0.5 * log2(e)
    attention_tcgen05_kernel<<<grid, 128, ATT_SMEM, st>>>(tm, (__nv_bfloat16*)d_out, T, d, scale_log2e);
    MW_LAUNCH_CHECK();
    return MW_OK;
}

}  // namespace mw

extern "C" mw_status mw_attention_bf16(const void* d_qkv, void* d_out, int B, int T, int n_heads, void* stream) {
    return mw::attention_launch(d_qkv, d_out, B, T, n_heads, (cudaStream_t)stream);
}
