// Persistent LSTM recurrence kernels for sm_100a (forward and BPTT), hidden size 1024.
//
// Reference semantics: torch.nn.LSTM as used by AR_Step (flowtron.py:654-655, 671-695): gate order
// i,f,g,o; c' = s(f) c + s(i) tanh(g); h' = s(o) tanh(c'); zero initial state; outputs are 0 at
// t >= len (pack_padded/pad_packed).  The input projection W_ih x + b_ih + b_hh is a plain GEMM done
// beforehand (gemm.cu); these kernels run the T-step dependency chain without returning to the host.
//
// Forward (lstm_fwd_kernel), one cooperative launch, 128 CTAs:
//   * CTA c owns hidden units [8c, 8c+8): its 32 rows of W_hh (4 gates x 8 units, fp16, 64 KB) are
//     TMA-loaded once and stay resident in shared memory as the UMMA B operand (N = 32).
//   * every step: h_{t-1} ([B,1024] fp16, read straight out of the layer's output tensor) is TMA-streamed
//     in 16 K-chunks into an mbarrier ring as the UMMA A operand (M = 128 rows, batch rows first; the
//     remaining rows are don't-care), tcgen05.mma accumulates the 32 gate pre-activations in TMEM,
//     the epilogue warps add the input projection, apply the LSTM cell in fp32 (cell state lives in
//     registers for the whole sequence), and publish h_t (fp16) + per-chunk release flags in global
//     memory; consumers acquire the flag, fence the async proxy, and TMA the chunk.
// Backward (lstm_bwd_kernel), 64 CTAs x 16 units: dh_{t-1} = dG_t W_hh with W_hh^T slice resident
//   (fp16, 128 KB), dG_t ([B,4096] fp16, loss-scaled) streamed through the ring, the pointwise LSTM backward in the
//   epilogue; dG is written once and reused by the wgrad/dgrad GEMMs.
#include "ptx.cuh"
#include "ft_internal.h"

namespace ft {

#define FT_TRACE(p, t, slot) do { if ((p).trace && blockIdx.x == 0) (p).trace[(t) * 8 + (slot)] = clock64(); } while (0)
static long long* g_lstm_trace = nullptr;      // set through ft_debug_set_lstm_trace
void set_lstm_trace(long long* p) { g_lstm_trace = p; }

constexpr int LH = 1024;
constexpr int LG = 4 * LH;
constexpr int KCH = 64;                        // K elements per 128-byte chunk (16-bit operands)
constexpr int LSTM_THREADS = 192;

// ------------------------------------------------------------------------------------------- forward
constexpr int FWD_CTAS = 128, FWD_UNITS = 8, FWD_N = 32, FWD_NCH = LH / KCH;     // 16 chunks / step
constexpr int FWD_W_BYTES = FWD_NCH * FWD_N * 128;                                 // 64 KB

struct LstmFwdParams {
    int T, B, Bbox, nslot;
    const float* xproj;        // [T*B, 4096] input projection + both biases
    const int* lens;           // [B] or null
    __half* hseq; long long ldh;   // [T*B, ldh] output (fp16), also the recurrent exchange buffer
    __half* gates;             // [T*B, 4096] post-activation i,f,g,o (fp16) or null
    float* cstate;             // [T*B, 1024] or null
    float* h32; long long ldh32;   // optional fp32 copy of h (or null)
    int* flags;                // [T * 16], zeroed by the launcher
    int* status;
    long long* trace;          // optional [T][8] clock64 stamps of CTA 0 (debug/profiling), or null
};

__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, LstmFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int slot_bytes = p.Bbox * 128;
    uint8_t* ring = smem;
    uint8_t* sW = smem + p.nslot * slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + FWD_W_BYTES);   // the M=128 over-read of the last ring slot lands in sW
    uint64_t* full = bars;                       // [nslot]
    uint64_t* empty = bars + 32;                 // [nslot]
    uint64_t* wbar = bars + 64;
    uint64_t* accum_full = bars + 65;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 66);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int nq = (p.B + 31) / 32;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmH);
        for (int s = 0; s < p.nslot; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(wbar, 1);
        mbar_init(accum_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // resident W_hh slice: chunk kc, gate g -> 8 rows x 128 B (one swizzle atom)
            mbar_expect_tx(wbar, FWD_W_BYTES);
            for (int kc = 0; kc < FWD_NCH; ++kc)
                for (int g = 0; g < 4; ++g)
                    tma_load_2d(sW + kc * (FWD_N * 128) + g * 1024, &tmW, wbar, kc * KCH, g * LH + FWD_UNITS * cta);
        }
        const int target = 8 * nq;                 // 8 producer CTAs per 64-unit chunk, one release per active quadrant warp
        int it = 0;
        for (int t = 1; t < p.T; ++t) {
            // all 16 chunk flags of step t-1 are polled IN PARALLEL (one lane each): a serial poll costs one L2
            // round trip per chunk and was 10 us/step in the first version
            if (lane < FWD_NCH) wait_flag_ge(&p.flags[(t - 1) * FWD_NCH + lane], target, p.status, 202);
            __syncwarp();
            if (lane == 0) {
                FT_TRACE(p, t, 0);                 // flags of step t-1 all visible
                fence_proxy_async();               // generic-proxy writes of other SMs -> async-proxy (TMA) reads
                for (int kc = 0; kc < FWD_NCH; ++kc, ++it) {
                    const int s = it % p.nslot, ph = (it / p.nslot) & 1;
                    mbar_wait(&empty[s], ph ^ 1, p.status, 201);
                    mbar_expect_tx(&full[s], slot_bytes);
                    tma_load_2d(ring + s * slot_bytes, &tmH, &full[s], kc * KCH, (t - 1) * p.B);
                }
                FT_TRACE(p, t, 1);                 // all TMA loads issued
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(wbar, 0, p.status, 203);
            const uint32_t idesc = umma_idesc(128, FWD_N, FMT_F16, FMT_F16, 0, 0);
            const uint32_t w0 = smem_u32(sW), r0 = smem_u32(ring);
            int it = 0;
            for (int t = 1; t < p.T; ++t) {
                for (int kc = 0; kc < FWD_NCH; ++kc, ++it) {
                    const int s = it % p.nslot, ph = (it / p.nslot) & 1;
                    mbar_wait(&full[s], ph, p.status, 204);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < KCH / 16; ++k) {
                        const uint64_t da = umma_smem_desc(r0 + s * slot_bytes + k * 32, 16, 1024);
                        const uint64_t db = umma_smem_desc(w0 + kc * (FWD_N * 128) + k * 32, 16, 1024);
                        umma_f16(tmem_base, da, db, idesc, (kc | k) != 0);
                    }
                    umma_commit(&empty[s]);
                    if (kc == 0) FT_TRACE(p, t, 2);          // first chunk landed
                }
                FT_TRACE(p, t, 3);                 // last chunk landed, all MMAs issued
                umma_commit(accum_full);
            }
        }
    } else {
        const int q = warp & 3;
        if (q < nq) {
            const int b = q * 32 + lane;
            const bool row_ok = b < p.B;
            const int len = (row_ok && p.lens) ? p.lens[b] : p.T;
            const int u0 = FWD_UNITS * cta;
            float c[FWD_UNITS];
#pragma unroll
            for (int j = 0; j < FWD_UNITS; ++j) c[j] = 0.f;
            for (int t = 0; t < p.T; ++t) {
                const long long r = static_cast<long long>(t) * p.B + b;
                float x[32];
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4* src = reinterpret_cast<const float4*>(p.xproj + r * LG + g * LH + u0);
                        const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
                        x[g * 8 + 0] = v0.x; x[g * 8 + 1] = v0.y; x[g * 8 + 2] = v0.z; x[g * 8 + 3] = v0.w;
                        x[g * 8 + 4] = v1.x; x[g * 8 + 5] = v1.y; x[g * 8 + 6] = v1.z; x[g * 8 + 7] = v1.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = 0.f;
                }
                if (t > 0) {
                    float acc[32];
                    mbar_wait(accum_full, (t - 1) & 1, p.status, 205);
                    tc_fence_after();
                    if (lane == 0 && q == 0) FT_TRACE(p, t, 4);      // accumulator complete
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16), acc);
                    tmem_ld_wait();
                    tc_fence_before();
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] += acc[j];
                }
                if (row_ok) {
                    const bool valid = t < len;
                    float hv[FWD_UNITS], gi[FWD_UNITS], gf[FWD_UNITS], gg[FWD_UNITS], go[FWD_UNITS];
#pragma unroll
                    for (int j = 0; j < FWD_UNITS; ++j) {
                        gi[j] = sigmoid_f(x[j]);
                        gf[j] = sigmoid_f(x[8 + j]);
                        gg[j] = tanh_f(x[16 + j]);
                        go[j] = sigmoid_f(x[24 + j]);
                        c[j] = gf[j] * c[j] + gi[j] * gg[j];
                        hv[j] = valid ? go[j] * tanh_f(c[j]) : 0.f;
                    }
                    {   // h_t (fp16) -> layer output / exchange buffer
                        __half2 h2[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(hv[2 * j], hv[2 * j + 1]);
                        *reinterpret_cast<uint4*>(p.hseq + r * p.ldh + u0) = *reinterpret_cast<uint4*>(h2);
                    }
                    if (p.h32) {
                        float4* d = reinterpret_cast<float4*>(p.h32 + r * p.ldh32 + u0);
                        d[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                        d[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                    }
                    if (p.gates) {
                        const float* gsrc[4] = {gi, gf, gg, go};
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            __half2 h2[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(gsrc[g][2 * j], gsrc[g][2 * j + 1]);
                            *reinterpret_cast<uint4*>(p.gates + r * LG + g * LH + u0) = *reinterpret_cast<uint4*>(h2);
                        }
                    }
                    if (p.cstate) {
                        float4* d = reinterpret_cast<float4*>(p.cstate + r * LH + u0);
                        d[0] = make_float4(c[0], c[1], c[2], c[3]);
                        d[1] = make_float4(c[4], c[5], c[6], c[7]);
                    }
                }
                __syncwarp();                      // lanes' stores happen-before lane 0's cumulative fence + release
                if (lane == 0) {
                    if (q == 0) FT_TRACE(p, t, 5);  // cell + stores done
                    __threadfence();
                    if (q == 0) FT_TRACE(p, t, 6);  // fence done
                    red_release_add(&p.flags[t * FWD_NCH + cta / 8], 1);
                    if (q == 0) FT_TRACE(p, t, 7);  // release issued
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<32>(tmem_base);
}

// ------------------------------------------------------------------------------------------- backward
constexpr int BWD_CTAS = 64, BWD_UNITS = 16, BWD_NCH = LG / KCH;                   // 64 chunks / step
constexpr int BWD_W_BYTES = BWD_NCH * BWD_UNITS * 128;                              // 128 KB

struct LstmBwdParams {
    int T, B, Bbox, nslot;
    const float* dh_ext; long long ldd;   // [T*B, ldd] gradient w.r.t. the layer outputs (fp32)
    const __half* gates;       // [T*B, 4096] saved i,f,g,o
    const float* cstate;       // [T*B, 1024] saved c_t
    const int* lens;
    __half* dG;                // [T*B, 4096] out: (loss-scaled) gradient w.r.t. gate pre-activations (fp16, saturating)
    int* flags;                // [T * 64]
    int* status;
    long long* trace;
};

__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_bwd_kernel(const __grid_constant__ CUtensorMap tmWT, const __grid_constant__ CUtensorMap tmG, LstmBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int slot_bytes = p.Bbox * 128;
    uint8_t* ring = smem;
    uint8_t* sW = smem + p.nslot * slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + BWD_W_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + 32;
    uint64_t* wbar = bars + 64;
    uint64_t* accum_full = bars + 65;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 66);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int nq = (p.B + 31) / 32;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmWT);
        tma_prefetch_desc(&tmG);
        for (int s = 0; s < p.nslot; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(wbar, 1);
        mbar_init(accum_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, BWD_W_BYTES);
            for (int kc = 0; kc < BWD_NCH; ++kc)       // W_hh^T rows [16c,16c+16), K chunk kc
                tma_load_2d(sW + kc * (BWD_UNITS * 128), &tmWT, wbar, kc * KCH, BWD_UNITS * cta);
        }
        const int target = 4 * nq;                     // 4 producer CTAs per 64-column chunk of dG
        int it = 0;
        for (int t = p.T - 2; t >= 0; --t) {
            for (int c = lane; c < BWD_NCH; c += 32)   // parallel poll of the 64 chunk flags of step t+1
                wait_flag_ge(&p.flags[(t + 1) * BWD_NCH + c], target, p.status, 212);
            __syncwarp();
            if (lane == 0) {
                fence_proxy_async();
                for (int kc = 0; kc < BWD_NCH; ++kc, ++it) {
                    const int s = it % p.nslot, ph = (it / p.nslot) & 1;
                    mbar_wait(&empty[s], ph ^ 1, p.status, 211);
                    mbar_expect_tx(&full[s], slot_bytes);
                    tma_load_2d(ring + s * slot_bytes, &tmG, &full[s], kc * KCH, (t + 1) * p.B);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(wbar, 0, p.status, 213);
            const uint32_t idesc = umma_idesc(128, BWD_UNITS, FMT_F16, FMT_F16, 0, 0);
            const uint32_t w0 = smem_u32(sW), r0 = smem_u32(ring);
            int it = 0;
            for (int t = p.T - 2; t >= 0; --t) {
                for (int kc = 0; kc < BWD_NCH; ++kc, ++it) {
                    const int s = it % p.nslot, ph = (it / p.nslot) & 1;
                    mbar_wait(&full[s], ph, p.status, 214);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < KCH / 16; ++k) {
                        const uint64_t da = umma_smem_desc(r0 + s * slot_bytes + k * 32, 16, 1024);
                        const uint64_t db = umma_smem_desc(w0 + kc * (BWD_UNITS * 128) + k * 32, 16, 1024);
                        umma_f16(tmem_base, da, db, idesc, (kc | k) != 0);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(accum_full);
            }
        }
    } else {
        const int q = warp & 3;
        if (q < nq) {
            const int b = q * 32 + lane;
            const bool row_ok = b < p.B;
            const int len = (row_ok && p.lens) ? p.lens[b] : p.T;
            const int u0 = BWD_UNITS * cta;
            float dcs[BWD_UNITS];
#pragma unroll
            for (int j = 0; j < BWD_UNITS; ++j) dcs[j] = 0.f;
            int step = 0;
            for (int t = p.T - 1; t >= 0; --t, ++step) {
                const long long r = static_cast<long long>(t) * p.B + b;
                float dh[BWD_UNITS], ct[BWD_UNITS], cp[BWD_UNITS];
                __half2 gt[4][BWD_UNITS / 2];
                const bool valid = row_ok && (t < len);
                if (valid) {
#pragma unroll
                    for (int j = 0; j < BWD_UNITS / 4; ++j) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(p.dh_ext + r * p.ldd + u0) + j);
                        dh[4 * j] = v.x; dh[4 * j + 1] = v.y; dh[4 * j + 2] = v.z; dh[4 * j + 3] = v.w;
                        const float4 cc = __ldg(reinterpret_cast<const float4*>(p.cstate + r * LH + u0) + j);
                        ct[4 * j] = cc.x; ct[4 * j + 1] = cc.y; ct[4 * j + 2] = cc.z; ct[4 * j + 3] = cc.w;
                        float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (t > 0) pp = __ldg(reinterpret_cast<const float4*>(p.cstate + (r - p.B) * LH + u0) + j);
                        cp[4 * j] = pp.x; cp[4 * j + 1] = pp.y; cp[4 * j + 2] = pp.z; cp[4 * j + 3] = pp.w;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint4* src = reinterpret_cast<const uint4*>(p.gates + r * LG + g * LH + u0);
                        uint4 a = __ldg(src), bb = __ldg(src + 1);
                        *reinterpret_cast<uint4*>(&gt[g][0]) = a;
                        *reinterpret_cast<uint4*>(&gt[g][4]) = bb;
                    }
                }
                float acc[BWD_UNITS];
#pragma unroll
                for (int j = 0; j < BWD_UNITS; ++j) acc[j] = 0.f;
                if (step > 0) {
                    mbar_wait(accum_full, (step - 1) & 1, p.status, 215);
                    tc_fence_after();
                    tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16), acc);
                    tmem_ld_wait();
                    tc_fence_before();
                }
                if (row_ok) {
                    __half2 out[4][BWD_UNITS / 2];
                    if (valid) {
#pragma unroll
                        for (int j2 = 0; j2 < BWD_UNITS / 2; ++j2) {
                            float da[4][2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int j = 2 * j2 + e;
                                const float gi = e ? __high2float(gt[0][j2]) : __low2float(gt[0][j2]);
                                const float gf = e ? __high2float(gt[1][j2]) : __low2float(gt[1][j2]);
                                const float gg = e ? __high2float(gt[2][j2]) : __low2float(gt[2][j2]);
                                const float go = e ? __high2float(gt[3][j2]) : __low2float(gt[3][j2]);
                                const float dht = dh[j] + acc[j];
                                const float tc = tanh_f(ct[j]);
                                const float dc = dcs[j] + dht * go * (1.f - tc * tc);
                                da[3][e] = dht * tc * go * (1.f - go);
                                da[0][e] = dc * gg * gi * (1.f - gi);
                                da[2][e] = dc * gi * (1.f - gg * gg);
                                da[1][e] = dc * cp[j] * gf * (1.f - gf);
                                dcs[j] = dc * gf;
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                out[g][j2] = __floats2half2_rn(fminf(fmaxf(da[g][0], -65504.f), 65504.f), fminf(fmaxf(da[g][1], -65504.f), 65504.f));
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int j2 = 0; j2 < BWD_UNITS / 2; ++j2) out[g][j2] = __floats2half2_rn(0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < BWD_UNITS; ++j) dcs[j] = 0.f;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4* dst = reinterpret_cast<uint4*>(p.dG + r * LG + g * LH + u0);
                        dst[0] = *reinterpret_cast<uint4*>(&out[g][0]);
                        dst[1] = *reinterpret_cast<uint4*>(&out[g][4]);
                    }
                }
                __syncwarp();
                if (lane < 4) {
                    __threadfence();
                    red_release_add(&p.flags[t * BWD_NCH + lane * 16 + cta / 4], 1);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<32>(tmem_base);
}

// ------------------------------------------------------------------------------------------- host
static int smem_optin() {
    static int v = -1;
    if (v < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    return v;
}

int launch_lstm_fwd(int T, int B, const float* xproj, const void* whh16, const int* lens, void* hseq16, long long ldh,
                    void* gates16, float* cstate, float* h32, long long ldh32, int* flags, cudaStream_t st) {
    if (T <= 0 || B <= 0) return 0;
    if (B > 128) return ft_set_error("lstm_fwd: batch > 128 per call not supported");
    LstmFwdParams p;
    p.T = T; p.B = B; p.Bbox = (B + 7) & ~7;
    const int slot = p.Bbox * 128;
    const int fixed = FWD_W_BYTES + 1024 + 1024;                  // W + barriers + align
    int nslot = (smem_optin() - fixed) / slot;
    if (nslot > 16) nslot = 16;
    if (nslot < 2) return ft_set_error("lstm_fwd: not enough shared memory for the h ring");
    p.nslot = nslot;
    p.xproj = xproj; p.lens = lens; p.hseq = static_cast<__half*>(hseq16); p.ldh = ldh;
    p.gates = static_cast<__half*>(gates16); p.cstate = cstate; p.h32 = h32; p.ldh32 = ldh32;
    p.flags = flags; p.status = ft_status_word(); p.trace = g_lstm_trace;
    CUtensorMap tmW, tmH;
    if (make_tmap_2d(&tmW, whh16, FMT_F16, LG, LH, LH, KCH, FWD_UNITS)) return -1;
    if (make_tmap_2d(&tmH, hseq16, FMT_F16, static_cast<long long>(T) * B, LH, ldh, KCH, p.Bbox)) return -1;
    if (cudaMemsetAsync(flags, 0, sizeof(int) * T * FWD_NCH, st) != cudaSuccess) return ft_set_error("lstm_fwd: memset failed");
    const int smem = nslot * slot + fixed;
    cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    TimeScope ts("lstm_fwd", T, B, 0, st);
    void* args[] = {&tmW, &tmH, &p};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_fwd_kernel), dim3(FWD_CTAS),
                                                dim3(LSTM_THREADS), args, smem, st);
    if (e != cudaSuccess) return ft_set_error(cudaGetErrorString(e));
    ft_count_launch(1);
    return ft_check_launch("lstm_fwd_kernel");
}

int launch_lstm_bwd(int T, int B, const float* dh_ext, long long ldd, const void* whhT16, const void* gates16,
                    const float* cstate, const int* lens, void* dG16, int* flags, cudaStream_t st) {
    if (T <= 0 || B <= 0) return 0;
    if (B > 128) return ft_set_error("lstm_bwd: batch > 128 per call not supported");
    LstmBwdParams p;
    p.T = T; p.B = B; p.Bbox = (B + 7) & ~7;
    const int slot = p.Bbox * 128;
    const int fixed = BWD_W_BYTES + 1024 + 1024;
    int nslot = (smem_optin() - fixed) / slot;          // over-read of the last slot lands in sW (128 KB >= 16 KB)
    if (nslot > 32) nslot = 32;
    if (nslot < 2) return ft_set_error("lstm_bwd: not enough shared memory for the dG ring");
    p.nslot = nslot;
    p.dh_ext = dh_ext; p.ldd = ldd; p.gates = static_cast<const __half*>(gates16); p.cstate = cstate; p.lens = lens;
    p.dG = static_cast<__half*>(dG16); p.flags = flags; p.status = ft_status_word(); p.trace = nullptr;
    CUtensorMap tmWT, tmG;
    if (make_tmap_2d(&tmWT, whhT16, FMT_F16, LH, LG, LG, KCH, BWD_UNITS)) return -1;
    if (make_tmap_2d(&tmG, dG16, FMT_F16, static_cast<long long>(T) * B, LG, LG, KCH, p.Bbox)) return -1;
    if (cudaMemsetAsync(flags, 0, sizeof(int) * T * BWD_NCH, st) != cudaSuccess) return ft_set_error("lstm_bwd: memset failed");
    const int smem = nslot * slot + fixed;
    cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    TimeScope ts("lstm_bwd", T, B, 0, st);
    void* args[] = {&tmWT, &tmG, &p};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lstm_bwd_kernel), dim3(BWD_CTAS),
                                                dim3(LSTM_THREADS), args, smem, st);
    if (e != cudaSuccess) return ft_set_error(cudaGetErrorString(e));
    ft_count_launch(1);
    return ft_check_launch("lstm_bwd_kernel");
}

}  // namespace ft
