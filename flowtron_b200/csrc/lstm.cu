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
//   * every step: h_{t-1} ([B,1024] fp16, read straight out of the layer's output tensor) arrives through
//     TWO 3-D TMA loads (8 K-chunks each, box {64 elems, B rows, 8 chunks}: the smem image is chunk-major
//     SWIZZLE_128B K-major tiles, the UMMA A operand with M = 128 rows of which the first B are real),
//     tcgen05.mma accumulates the 32 gate pre-activations in TMEM, the accumulator rows are handed to all
//     four epilogue warps through shared memory, which add the input projection, apply the cell in fp32
//     (cell state lives in registers for the whole sequence), store h_t (fp16) and publish ONE release per
//     CTA into a per-(step, chunk) counter; consumers poll the 16 counters in parallel (one lane each),
//     fence the async proxy, and issue the next loads.  gates / c for the backward pass are stored AFTER
//     the release so they stay off the critical path.
// Backward (lstm_bwd_kernel), 64 CTAs x 16 units: dh_{t-1} = dG_t W_hh with the W_hh^T slice resident
//   (fp16, 128 KB), dG_t ([B,4096] fp16, loss-scaled) streamed in groups of chunks through a small ring,
//   the pointwise LSTM backward in the epilogue; dG is written once and reused by the wgrad/dgrad GEMMs.
//
// History (profiles/, tools/trace_lstm.py): v1 polled the chunk flags and issued 16 (64) 2-D TMA loads
// serially from one thread: 4 us of an 8.9 us forward step was TMA issue, 0.85 us a single-warp epilogue,
// 1 us fence+release behind 112 B/lane of stores.
#include <cstdlib>

#include "ptx.cuh"
#include "ft_internal.h"

namespace ft {

#define FT_TRACE(p, t, slot) do { if ((p).trace && blockIdx.x == 0) (p).trace[(t) * 16 + (slot)] = clock64(); } while (0)
static long long* g_lstm_trace = nullptr;      // set through ft_debug_set_lstm_trace
void set_lstm_trace(long long* p) { g_lstm_trace = p; }

constexpr int LH = 1024;
constexpr int LG = 4 * LH;
constexpr int KCH = 64;                        // K elements per 128-byte chunk (16-bit operands)
constexpr int LSTM_THREADS = 192;              // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
constexpr int EPI_THREADS = 128;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the 4 epilogue warps

// ------------------------------------------------------------------------------------------- forward
constexpr int FWD_NCH = LH / KCH;                  // 16 chunks / step
// FWD_UNITS hidden units per CTA: 8 -> 128 CTAs (whole GPU, lowest latency), 16 -> 64 CTAs (two such launches on two
// streams -- one per half batch -- run concurrently and hide each other's exchange latency)

struct LstmFwdParams {
    int T, B, Bbox;
    const float* xproj;        // [T*B, 4096] input projection + both biases (null when the projection is folded in)
    const float* b_ih; const float* b_hh;   // folded-projection form only: the two bias vectors [4096]
    const int* lens;           // [B] or null
    __half* hseq; long long ldh;   // [T*B, ldh] output (fp16), also the recurrent exchange buffer
    __half* gates;             // [T*B, 4096] post-activation i,f,g,o (fp16) or null
    float* cstate;             // [T*B, 1024] or null
    float* h32; long long ldh32;   // optional fp32 copy of h (or null)
    int* flags;                // [T * 16], zeroed by the launcher
    int* status;
    long long* trace;          // optional [T][8] clock64 stamps of CTA 0 (debug/profiling), or null
};

// ------------------------------------------------------------------------------------------- backward
constexpr int BWD_CTAS = 64, BWD_UNITS = 16, BWD_NCH = LG / KCH;                   // 64 chunks / step
constexpr int BWD_W_BYTES = BWD_NCH * BWD_UNITS * 128;                              // 128 KB

struct LstmBwdParams {
    int T, B, Bbox;
    int gs, ng, nring;         // chunks per TMA group, groups per step (64 / gs), ring depth in groups
    const float* dh_ext; long long ldd;   // [T*B, ldd] gradient w.r.t. the layer outputs (fp32, loss-scaled)
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
    const int group_bytes = p.gs * slot_bytes;
    uint8_t* ring = smem;                                        // [nring groups][gs chunks][Bbox rows][128 B]
    uint8_t* sW = smem + p.nring * group_bytes;                  // over-read of the last ring chunk lands here
    float* sAcc = reinterpret_cast<float*>(sW + BWD_W_BYTES);    // [128 rows][17]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sAcc + 128 * 17 + 1);
    bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bars) + 7) & ~uintptr_t(7));
    uint64_t* full = bars;                       // [nring]
    uint64_t* empty = bars + 8;                  // [nring]
    uint64_t* wbar = bars + 16;
    uint64_t* accum_full = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int nq = (p.B + 31) / 32;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmWT);
        tma_prefetch_desc(&tmG);
        for (int s = 0; s < p.nring; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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
        int s = 0, ph = 0;                             // ring slot / phase, advanced incrementally (no div/mod)
        for (int t = p.T - 2; t >= 0; --t) {
            for (int c = lane; c < BWD_NCH; c += 32)   // parallel poll: 4 producer CTAs per 64-column chunk of dG_{t+1}
                wait_flag_ge(&p.flags[(t + 1) * BWD_NCH + c], 4, p.status, 212);
            __syncwarp();
            if (elect_one()) { FT_TRACE(p, t, 0); fence_proxy_async_global(); }
            __syncwarp();
            for (int g = 0; g < p.ng; ++g) {
                mbar_wait(&empty[s], ph ^ 1, p.status, 211);
                if (elect_one()) {
                    mbar_expect_tx(&full[s], group_bytes);
                    tma_load_3d(ring + s * group_bytes, &tmG, &full[s], 0, (t + 1) * p.B, g * p.gs);
                    if (g == 0) FT_TRACE(p, t, 1);
                    if (g == p.ng - 1) FT_TRACE(p, t, 6);
                }
                __syncwarp();
                if (++s == p.nring) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        mbar_wait(wbar, 0, p.status, 213);
        const uint32_t idesc = umma_idesc(128, BWD_UNITS, FMT_F16, FMT_F16, 0, 0);
        const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t da_ring = umma_smem_desc(smem_u32(ring), 16, 1024), db_base = umma_smem_desc(smem_u32(sW), 16, 1024);
        const uint64_t a_chunk = static_cast<uint64_t>(slot_bytes >> 4), a_group = static_cast<uint64_t>(group_bytes >> 4);
        const uint64_t b_chunk = (BWD_UNITS * 128) >> 4;
        int s = 0, ph = 0;
        uint64_t da_slot = da_ring;
        for (int t = p.T - 2; t >= 0; --t) {
            uint64_t db = db_base;
            for (int g = 0; g < p.ng; ++g) {
                mbar_wait(&full[s], ph, p.status, 214);
                tc_fence_after();
                if (elect_one()) {
                    if (g == 0) FT_TRACE(p, t, 2);
                    uint64_t xa = da_slot, xb = db;
                    uint32_t acc = g != 0;
                    for (int c = 0; c < p.gs; ++c) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16(tmem_d, xa + 2 * k, xb + 2 * k, idesc, acc);
                            acc = 1;
                        }
                        xa += a_chunk;
                        xb += b_chunk;
                    }
                    umma_commit(&empty[s]);
                    if (g == p.ng - 1) { FT_TRACE(p, t, 3); umma_commit(accum_full); }
                }
                __syncwarp();
                db += static_cast<uint64_t>(p.gs) * b_chunk;
                da_slot += a_group;
                if (++s == p.nring) { s = 0; ph ^= 1; da_slot = da_ring; }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue: items (batch row, 4 units)
        const int q = warp & 3;
        const int et = threadIdx.x - 64;
        const int n_items = p.B * 4;
        const int u0 = BWD_UNITS * cta;
        float dcs[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) dcs[s][j] = 0.f;
        int len_i[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int item = et + s * EPI_THREADS;
            len_i[s] = (item < n_items && p.lens) ? p.lens[item >> 2] : p.T;
        }
        int step = 0;
        for (int t = p.T - 1; t >= 0; --t, ++step) {
            float dh[2][4], ct[2][4], cp[2][4];
            __half2 gt[2][4][2];                              // [item][gate][2 x half2]
            bool valid[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int item = et + s * EPI_THREADS;
                valid[s] = (item < n_items) && (t < len_i[s]);
                if (valid[s]) {
                    const int b = item >> 2, uq = item & 3;
                    const long long r = static_cast<long long>(t) * p.B + b;
                    const int uo = u0 + 4 * uq;
                    const float4 v = __ldg(reinterpret_cast<const float4*>(p.dh_ext + r * p.ldd + uo));
                    dh[s][0] = v.x; dh[s][1] = v.y; dh[s][2] = v.z; dh[s][3] = v.w;
                    const float4 cc = __ldg(reinterpret_cast<const float4*>(p.cstate + r * LH + uo));
                    ct[s][0] = cc.x; ct[s][1] = cc.y; ct[s][2] = cc.z; ct[s][3] = cc.w;
                    float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t > 0) pp = __ldg(reinterpret_cast<const float4*>(p.cstate + (r - p.B) * LH + uo));
                    cp[s][0] = pp.x; cp[s][1] = pp.y; cp[s][2] = pp.z; cp[s][3] = pp.w;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(p.gates + r * LG + g * LH + uo));
                        *reinterpret_cast<uint2*>(&gt[s][g][0]) = pk;
                    }
                }
            }
            if (step > 0) {
                if (q < nq) {
                    float acc[16];
                    mbar_wait(accum_full, (step - 1) & 1, p.status, 215);
                    tc_fence_after();
                    if (et == 64) FT_TRACE(p, t, 4);
                    tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16), acc);
                    tmem_ld_wait();
                    tc_fence_before();
                    float* dst = sAcc + (q * 32 + lane) * 17;
#pragma unroll
                    for (int j = 0; j < 16; ++j) dst[j] = acc[j];
                }
                epi_bar();
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int item = et + s * EPI_THREADS;
                if (item < n_items) {
                    const int b = item >> 2, uq = item & 3;
                    const long long r = static_cast<long long>(t) * p.B + b;
                    __half2 out[4][2];
                    if (valid[s]) {
#pragma unroll
                        for (int j2 = 0; j2 < 2; ++j2) {
                            float da[4][2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int j = 2 * j2 + e;
                                const float gi = e ? __high2float(gt[s][0][j2]) : __low2float(gt[s][0][j2]);
                                const float gf = e ? __high2float(gt[s][1][j2]) : __low2float(gt[s][1][j2]);
                                const float gg = e ? __high2float(gt[s][2][j2]) : __low2float(gt[s][2][j2]);
                                const float go = e ? __high2float(gt[s][3][j2]) : __low2float(gt[s][3][j2]);
                                float dht = dh[s][j];
                                if (step > 0) dht += sAcc[b * 17 + 4 * uq + j];
                                const float tc = tanh_f(ct[s][j]);
                                const float dc = dcs[s][j] + dht * go * (1.f - tc * tc);
                                da[3][e] = dht * tc * go * (1.f - go);
                                da[0][e] = dc * gg * gi * (1.f - gi);
                                da[2][e] = dc * gi * (1.f - gg * gg);
                                da[1][e] = dc * cp[s][j] * gf * (1.f - gf);
                                dcs[s][j] = dc * gf;
                            }
#pragma unroll
                            for (int g = 0; g < 4; ++g)
                                out[g][j2] = __floats2half2_rn(fminf(fmaxf(da[g][0], -65504.f), 65504.f),
                                                               fminf(fmaxf(da[g][1], -65504.f), 65504.f));
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) { out[g][0] = __floats2half2_rn(0.f, 0.f); out[g][1] = out[g][0]; }
#pragma unroll
                        for (int j = 0; j < 4; ++j) dcs[s][j] = 0.f;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        *reinterpret_cast<uint2*>(p.dG + r * LG + g * LH + u0 + 4 * uq) = *reinterpret_cast<uint2*>(&out[g][0]);
                }
            }
            epi_bar();                                        // every dG_t store of this CTA precedes the releases
            if (et == 0) FT_TRACE(p, t, 5);
            if (et < 4) red_release_add(&p.flags[t * BWD_NCH + et * 16 + cta / 4], 1);   // one release per gate chunk
            if (et == 0) FT_TRACE(p, t, 7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<32>(tmem_base);
}

// ------------------------------------------------------------------------------------------- backward, split-K cluster
// 64 CTAs in 16 clusters of 4.  Cluster c owns hidden units [64c, 64c+64); rank r reduces over gate r's quarter of
// the contraction (K = 1024 of 4096):  partial_r[b, u] = sum_{k in gate r} dG_{t+1}[b, k] W_hh[k, u].
//   * the rank's W_hh^T block [64 units x 1024] (fp16, 128 KB) is resident; per step it streams only ITS quarter of
//     dG_{t+1} (B x 1024 fp16 = 64 KB at B=32; the v2 kernel broadcast all 256 KB to every CTA, 6.3 us of an 8.7 us step);
//   * partial sums (fp32 [B, 64]) go to shared memory; a cluster-scope mbarrier (4 remote arrives) publishes them;
//     each rank then finishes 16 of the 64 units, reading the 3 other partials over DSMEM;
//   * pointwise LSTM backward + dG_t stores + per-chunk release flags exactly as in the single-CTA kernel.
constexpr int B4_CTAS = 64, B4_CLUSTER = 4, B4_UNITS = 64, B4_OWN = 16, B4_NCH = LH / KCH;     // 16 chunks / step / rank
constexpr int B4_W_BYTES = B4_NCH * B4_UNITS * 128;                                               // 128 KB
constexpr int B4_UP = 36;                                                                          // received-slice row pitch (floats): 32 batch columns, 16-byte aligned rows
constexpr int B4_XCHG_FLOATS = 2 * 4 * 16 * B4_UP;                                                 // [2 parities][4 source ranks][16 units][B4_UP]

// The kernel runs steps [t0, t1) of the sequence, highest step first: the whole sequence in one launch (t0 = 0, t1 = T), or
// one chunk of the layer pipeline (ar_step.cu), resuming dc*f from `dc_carry` ([B,1024] fp32, written by the chunk above) and
// dG_{t1} from the dG tensor.  Flags are indexed relative to t0, mbarrier phases are counted over the steps that have a
// recurrent term.
struct LstmBwdChunkParams : LstmBwdParams {
    int t0, t1;                // steps [t0, t1); flags: [(t1 - t0) * 64] ints, zeroed by the launcher
    float* dc_carry;           // [B, 1024]
};

__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_bwd4_kernel(const __grid_constant__ CUtensorMap tmWT, const __grid_constant__ CUtensorMap tmG, LstmBwdChunkParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int slot_bytes = p.Bbox * 128;
    uint8_t* sA = smem;                                          // [16 chunks][Bbox rows][128 B]
    uint8_t* sW = smem + B4_NCH * slot_bytes;                    // [16 chunks][64 rows][128 B]
    float* sPart = reinterpret_cast<float*>(sW + B4_W_BYTES);    // received partial sums: [2 parities][4 source ranks][16 units][B4_UP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sPart + B4_XCHG_FLOATS);
    uint64_t* full = bars;                       // [2]
    uint64_t* wbar = bars + 2;
    uint64_t* accum_full = bars + 3;
    uint64_t* part_bar = bars + 4;               // [2] : 4 cluster-scope arrivals each
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const int cl = blockIdx.x / B4_CLUSTER;
    constexpr int GS = 8, NG = B4_NCH / GS;
    const bool top = p.t1 == p.T;                                // the chunk that holds the last time step
    const int t_hi = top ? p.T - 2 : p.t1 - 1;                   // highest step that has a recurrent term (dG_{t+1} exists)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmWT);
        tma_prefetch_desc(&tmG);
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_init(wbar, 1);
        mbar_init(accum_full, 1);
        mbar_init(&part_bar[0], B4_CLUSTER); mbar_init(&part_bar[1], B4_CLUSTER);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<32>(tmem_slot);                    // accumulator: 64 lanes (units) x Bbox <= 32 columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();                                          // every rank's barriers exist before any remote arrive

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(wbar, B4_W_BYTES);
            for (int kc = 0; kc < B4_NCH; ++kc)                  // W_hh^T rows [64c, 64c+64), columns of gate `rank`
                tma_load_2d(sW + kc * (B4_UNITS * 128), &tmWT, wbar, rank * LH + kc * KCH, B4_UNITS * cl);
        }
        for (int t = t_hi; t >= p.t0; --t) {
            // gate `rank` of dG_{t+1}: 16 chunks, each released by the 4 ranks of one cluster.  dG_{t1} (first step of a
            // lower chunk) was written by the previous launch on this stream: nothing to wait for.
            if (t + 1 < p.t1 && lane < B4_NCH)
                wait_flag_ge(&p.flags[(t + 1 - p.t0) * BWD_NCH + rank * B4_NCH + lane], 4, p.status, 232);
            __syncwarp();
            if (elect_one()) {
                FT_TRACE(p, t, 0);
                fence_proxy_async_global();
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    mbar_expect_tx(&full[g], GS * slot_bytes);
                    tma_load_3d(sA + g * GS * slot_bytes, &tmG, &full[g], 0, (t + 1) * p.B, rank * B4_NCH + g * GS);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        mbar_wait(wbar, 0, p.status, 233);
        // swapped operands (see lstm_fwd_kernel): A = resident W_hh^T block (M = 64 units), B = dG_{t+1} quarter (N = Bbox rows)
        const uint32_t idesc = umma_idesc(64, p.Bbox, FMT_F16, FMT_F16, 0, 0);
        const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t da_base = umma_smem_desc(smem_u32(sW), 16, 1024), db_base = umma_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t a_chunk = (B4_UNITS * 128) >> 4, b_chunk = static_cast<uint64_t>(slot_bytes >> 4);
        int step = 0;
        for (int t = t_hi; t >= p.t0; --t, ++step) {
            const int ph = step & 1;
            uint64_t da = da_base, db = db_base;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                mbar_wait(&full[g], ph, p.status, 234);
                tc_fence_after();
                if (elect_one()) {
                    if (g == 0) FT_TRACE(p, t, 2);
                    uint64_t xa = da, xb = db;
#pragma unroll
                    for (int c = 0; c < GS; ++c) {
#pragma unroll
                        for (int k = 0; k < KCH / 16; ++k)
                            umma_f16(tmem_d, xa + 2 * k, xb + 2 * k, idesc, (g | c | k) != 0);
                        xa += a_chunk;
                        xb += b_chunk;
                    }
                    if (g == NG - 1) { FT_TRACE(p, t, 3); umma_commit(accum_full); }
                }
                __syncwarp();
                da += GS * a_chunk;
                db += GS * b_chunk;
            }
        }
    } else {
        const int q = warp & 3;
        const int et = threadIdx.x - 64;
        const int n_items = p.B * 4;                              // (batch row, quad of units) over this rank's 16 final units
        const int u0 = B4_UNITS * cl + B4_OWN * rank;
        const int item = et;
        const bool has_item = item < n_items;
        const int ib = item >> 2, uq = item & 3;
        float dcs[4] = {0.f, 0.f, 0.f, 0.f};
        const int len = (has_item && p.lens) ? p.lens[ib] : p.T;
        if (!top && has_item) {                                   // resume dc * f carried out of the chunk above
            const float4 cv = *reinterpret_cast<const float4*>(p.dc_carry + static_cast<long long>(ib) * LH + B4_UNITS * cl + B4_OWN * rank + 4 * uq);
            dcs[0] = cv.x; dcs[1] = cv.y; dcs[2] = cv.z; dcs[3] = cv.w;
        }
        int step = 0;
        for (int t = p.t1 - 1; t >= p.t0; --t, ++step) {
            const int rs = top ? step - 1 : step;                 // index of this step among the steps with a recurrent term
            float dh[4], ct[4], cp[4];
            __half2 gt[4][2];
            const bool valid = has_item && (t < len);
            const long long r = static_cast<long long>(t) * p.B + ib;
            const int uo = u0 + 4 * uq;
            if (valid) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.dh_ext + r * p.ldd + uo));
                dh[0] = v.x; dh[1] = v.y; dh[2] = v.z; dh[3] = v.w;
                const float4 cc = __ldg(reinterpret_cast<const float4*>(p.cstate + r * LH + uo));
                ct[0] = cc.x; ct[1] = cc.y; ct[2] = cc.z; ct[3] = cc.w;
                float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t > 0) pp = __ldg(reinterpret_cast<const float4*>(p.cstate + (r - p.B) * LH + uo));
                cp[0] = pp.x; cp[1] = pp.y; cp[2] = pp.z; cp[3] = pp.w;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(p.gates + r * LG + g * LH + uo));
                    *reinterpret_cast<uint2*>(&gt[g][0]) = pk;
                }
            }
            float rec[4] = {0.f, 0.f, 0.f, 0.f};                 // recurrent part of dh_t: sum over the 4 ranks' partials
            if (rs >= 0) {
                const int par = rs & 1;
                // Accumulator rows = the cluster block's 64 units: quadrant q (= this warp) holds units [16q, 16q+16), exactly the
                // units rank q finishes.  PUSH: each warp sends its 16 x Bbox partial sums straight into rank q's shared memory
                // (slot of THIS source rank, [unit][batch] rows of 32 floats), then one release-arrive per rank publishes them
                // and every rank sums four LOCAL slices -- one DSMEM hop on the critical path (r1 pulled: arrive, then a remote
                // load round trip), and all four epilogue warps move data instead of one.
                mbar_wait(accum_full, rs & 1, p.status, 235);
                tc_fence_after();
                if (et == 64) FT_TRACE(p, t, 4);
                {
                    float acc[32];
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16), acc);
                    tmem_ld_wait();
                    if (lane < 16) {
                        const uint32_t dst = mapa_shared(smem_u32(sPart) + static_cast<uint32_t>((((par * 4 + rank) * 16 + lane) * B4_UP) * 4), q);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) st_dsmem_f4(dst + j * 4, acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                    }
                }
                tc_fence_before();
                if (et == 64) FT_TRACE(p, t, 6);
                epi_bar();
                // publish: one release-arrive on every rank's barrier, issued by 4 different threads so the four
                // cluster-scope releases overlap (issued serially by one thread they cost ~1.8 us per step)
                if (et < B4_CLUSTER) mbar_arrive_cluster(mapa_shared(smem_u32(&part_bar[par]), et));
                mbar_wait_cluster(&part_bar[par], (rs >> 1) & 1, p.status, 236);
                if (et == 0) FT_TRACE(p, t, 1);               // all partials visible
                if (has_item) {
#pragma unroll
                    for (int rr = 0; rr < B4_CLUSTER; ++rr) {
                        const float* src = sPart + ((par * 4 + rr) * 16 + 4 * uq) * B4_UP + ib;
#pragma unroll
                        for (int j = 0; j < 4; ++j) rec[j] += src[j * B4_UP];
                    }
                }
            }
            if (has_item) {
                __half2 out[4][2];
                if (valid) {
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        float da[4][2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int j = 2 * j2 + e;
                            const float gi = e ? __high2float(gt[0][j2]) : __low2float(gt[0][j2]);
                            const float gf = e ? __high2float(gt[1][j2]) : __low2float(gt[1][j2]);
                            const float gg = e ? __high2float(gt[2][j2]) : __low2float(gt[2][j2]);
                            const float go = e ? __high2float(gt[3][j2]) : __low2float(gt[3][j2]);
                            const float dht = dh[j] + rec[j];
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
                            out[g][j2] = __floats2half2_rn(fminf(fmaxf(da[g][0], -65504.f), 65504.f),
                                                           fminf(fmaxf(da[g][1], -65504.f), 65504.f));
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) { out[g][0] = __floats2half2_rn(0.f, 0.f); out[g][1] = out[g][0]; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) dcs[j] = 0.f;
                }
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<uint2*>(p.dG + r * LG + g * LH + uo) = *reinterpret_cast<uint2*>(&out[g][0]);
            }
            epi_bar();                                            // every dG_t store of this CTA precedes the releases
            if (et == 0) FT_TRACE(p, t, 5);
            if (et < 4) red_release_add(&p.flags[(t - p.t0) * BWD_NCH + et * 16 + cl], 1);     // chunk (gate et, units 64cl..) : 4 ranks
            if (et == 0) FT_TRACE(p, t, 7);
        }
        if (p.t0 > 0 && has_item)                                 // hand dc * f to the chunk below
            *reinterpret_cast<float4*>(p.dc_carry + static_cast<long long>(ib) * LH + B4_UNITS * cl + B4_OWN * rank + 4 * uq) =
                make_float4(dcs[0], dcs[1], dcs[2], dcs[3]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                          // nobody exits while a peer may still read its partials
    if (warp == 1) tmem_dealloc<32>(tmem_base);
}


// ------------------------------------------------------------------------------------------- forward kernel
// Steps [t0, t1) of the recurrence: the whole sequence in one launch (t0 = 0, t1 = T), or one chunk of the layer pipeline
// (ar_step.cu: two layers run as two 64-CTA kernels on two streams, chunk by chunk, one chunk apart), resuming h from the
// output tensor (row t0-1) and c from the saved cell states.  Flags are indexed relative to t0, mbarrier phases relative to
// the first step that has a recurrent term.
struct LstmFwdChunkParams : LstmFwdParams {
    int t0, t1;                // steps [t0, t1); flags: [(t1 - t0) * 16] ints, zeroed by the launcher
};

// kXIn: the layer's INPUT projection is folded into the recurrent contraction (narrow inputs: the attention LSTM's 80 mel
// channels).  The step's input row block x_t ([B, <=128] fp16, zero-filled by TMA beyond the real width) is loaded next to
// h_{t-1} as XCH more K chunks, W_ih's slice sits next to W_hh's in shared memory, and b_ih + b_hh are added in the epilogue:
// no [T*B, 4096] fp32 projection tensor is written (0.5 ms GEMM + 524 MB write + 524 MB read per flow at B=32, T=1000).
// Requires x_0 == 0 (the teacher-forcing shift guarantees it): step 0 has no contraction at all.
constexpr int XCH = 2;                                           // extra K chunks (128 input channels max)

template <int FWD_GS, int FWD_UNITS, bool kXIn>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWx,
                const __grid_constant__ CUtensorMap tmX, LstmFwdChunkParams p) {
    constexpr int FWD_NG = FWD_NCH / FWD_GS;                     // TMA groups per step
    constexpr int NCHT = FWD_NCH + (kXIn ? XCH : 0);             // K chunks per step: h (16) [+ x (2)]
    constexpr int A_SLOTS = FWD_NCH + (kXIn ? 2 * XCH : 0);      // x_t is double-buffered: it is prefetched one step ahead
    constexpr int FWD_N = 4 * FWD_UNITS, FWD_W_BYTES = NCHT * FWD_N * 128, FWD_PROD = KCH / FWD_UNITS;
    constexpr int AP = FWD_N + 1;                                // accumulator staging pitch
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int slot_bytes = p.Bbox * 128;
    uint8_t* sA = smem;                                          // [16 (+2x2) chunks][Bbox rows][128 B]: h chunks, then two x buffers
    uint8_t* sW = smem + A_SLOTS * slot_bytes;
    float* sAcc = reinterpret_cast<float*>(sW + FWD_W_BYTES);    // [32 * nq rows][AP]
    const int nq = (p.B + 31) / 32;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sAcc + ((nq * 32 * AP + 1) & ~1));
    uint64_t* full = bars;                       // [FWD_NG] (<= 16)
    uint64_t* wbar = bars + 16;
    uint64_t* accum_full = bars + 17;
    uint64_t* xfull = bars + 18;                 // [2] (kXIn)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int tm0 = p.t0 > 1 ? p.t0 : 1;                         // first step with a recurrent term

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmH);
        if (kXIn) { tma_prefetch_desc(&tmWx); tma_prefetch_desc(&tmX); }
        for (int g = 0; g < FWD_NG; ++g) mbar_init(&full[g], 1);
        mbar_init(wbar, 1);
        mbar_init(accum_full, 1);
        mbar_init(&xfull[0], 1); mbar_init(&xfull[1], 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<64>(tmem_slot);             // accumulator: 64 lanes (W rows) x Bbox <= 64 columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // resident W_hh slice: chunk kc, gate g -> FWD_UNITS rows x 128 B
            mbar_expect_tx(wbar, FWD_W_BYTES);
            for (int kc = 0; kc < FWD_NCH; ++kc)
                for (int g = 0; g < 4; ++g)
                    tma_load_2d(sW + kc * (FWD_N * 128) + g * (FWD_UNITS * 128), &tmW, wbar, kc * KCH, g * LH + FWD_UNITS * cta);
            if (kXIn)                                          // W_ih slice behind it: chunks 16, 17 (columns beyond the input width: 0)
                for (int kc = 0; kc < XCH; ++kc)
                    for (int g = 0; g < 4; ++g)
                        tma_load_2d(sW + (FWD_NCH + kc) * (FWD_N * 128) + g * (FWD_UNITS * 128), &tmWx, wbar, kc * KCH, g * LH + FWD_UNITS * cta);
        }
        if (kXIn && tm0 < p.t1 && elect_one()) {               // x of the first contraction step: no dependency, load it now
            mbar_expect_tx(&xfull[0], XCH * slot_bytes);
#pragma unroll
            for (int kc = 0; kc < XCH; ++kc) tma_load_2d(sA + (FWD_NCH + kc) * slot_bytes, &tmX, &xfull[0], kc * KCH, tm0 * p.B);
        }
        __syncwarp();
        for (int t = tm0; t < p.t1; ++t) {
            // the first step of a chunk has nothing to wait for: h_{t0-1} was written by the previous launch on this stream
            // (kernel boundary) and the A buffer is untouched
            if (t > p.t0 && lane < FWD_NCH) wait_flag_ge(&p.flags[(t - 1 - p.t0) * FWD_NCH + lane], FWD_PROD, p.status, 212);
            __syncwarp();
            if (elect_one()) {
                FT_TRACE(p, t, 0);
                fence_proxy_async_global();        // generic-proxy writes of other SMs -> async-proxy (TMA) reads
                FT_TRACE(p, t, 6);
#pragma unroll
                for (int g = 0; g < FWD_NG; ++g) {
                    mbar_expect_tx(&full[g], FWD_GS * slot_bytes);
                    tma_load_3d(sA + g * FWD_GS * slot_bytes, &tmH, &full[g], 0, (t - 1) * p.B, g * FWD_GS);
                }
                if (kXIn && t + 1 < p.t1) {
                    // x_{t+1} does not depend on the recurrence: prefetch it one step ahead (it comes from HBM, not from L2
                    // like h: riding on this step's barrier cost 0.6 us per step).  Its buffer was last read by step t-1's MMAs,
                    // which have retired (all flags of step t-1, ours included, were seen above).
                    const int xs = (t + 1 - tm0) & 1;
                    mbar_expect_tx(&xfull[xs], XCH * slot_bytes);
#pragma unroll
                    for (int kc = 0; kc < XCH; ++kc)
                        tma_load_2d(sA + (FWD_NCH + xs * XCH + kc) * slot_bytes, &tmX, &xfull[xs], kc * KCH, (t + 1) * p.B);
                }
                FT_TRACE(p, t, 1);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // warp-uniform loop, one elected lane issues (see gemm.cu: descriptors must be uniform-register operands)
        mbar_wait(wbar, 0, p.status, 213);
        // SWAPPED operands (r2): A = the resident W slice (M = 64 accumulator rows: the CTA's 4*UNITS gate rows, plus -- for
        // UNITS = 8 -- 32 don't-care rows read from the next chunk), B = h_{t-1} (N = Bbox batch rows, exactly what TMA
        // delivered).  The r1 form (A = h padded to M = 128, B = W) made every instruction read 128 rows x 32 B of A from
        // shared memory, 3/4 of it padding: 5-6 KB per MMA against the SM's 128 B/clk = 40-48 clk per instruction instead of the
        // 16-clk tensor floor (trace: 64 MMAs took 1.2-1.5 us of a 4-5 us step).  Now 2 KB + 1 KB per instruction.
        const uint32_t idesc = umma_idesc(64, p.Bbox, FMT_F16, FMT_F16, 0, 0);
        const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_base, 0);
        const uint64_t da_base = umma_smem_desc(smem_u32(sW), 16, 1024), db_base = umma_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t a_chunk = (FWD_N * 128) >> 4, b_chunk = static_cast<uint64_t>(slot_bytes >> 4);
        for (int t = tm0; t < p.t1; ++t) {
            const int ph = (t - tm0) & 1;
            uint64_t da = da_base, db = db_base;
#pragma unroll
            if (kXIn) mbar_wait(&xfull[(t - tm0) & 1], ((t - tm0) >> 1) & 1, p.status, 216);
            for (int g = 0; g < FWD_NG; ++g) {
                mbar_wait(&full[g], ph, p.status, 214);
                tc_fence_after();
                if (elect_one()) {
                    if (g == 0) FT_TRACE(p, t, 2);
                    if (g == FWD_NG - 1) FT_TRACE(p, t, 5);
                    if (kXIn && g == 0) {                       // the (prefetched) input chunks open the accumulation
                        const int xs = (t - tm0) & 1;
                        uint64_t ya = da_base + FWD_NCH * a_chunk, yb = db_base + (FWD_NCH + xs * XCH) * b_chunk;
#pragma unroll
                        for (int c = 0; c < XCH; ++c) {
#pragma unroll
                            for (int k = 0; k < KCH / 16; ++k)
                                umma_f16(tmem_d, ya + 2 * k, yb + 2 * k, idesc, (c | k) != 0);
                            ya += a_chunk;
                            yb += b_chunk;
                        }
                    }
                    uint64_t xa = da, xb = db;
#pragma unroll
                    for (int c = 0; c < FWD_GS; ++c) {
#pragma unroll
                        for (int k = 0; k < KCH / 16; ++k)
                            umma_f16(tmem_d, xa + 2 * k, xb + 2 * k, idesc, kXIn || (g | c | k) != 0);
                        xa += a_chunk;
                        xb += b_chunk;
                    }
                    if (g == FWD_NG - 1) { FT_TRACE(p, t, 3); umma_commit(accum_full); }
                }
                __syncwarp();
                da += FWD_GS * a_chunk;
                db += FWD_GS * b_chunk;
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue: 4 warps, 128 threads
        const int q = warp & 3;                               // TMEM lane quadrant this warp may read
        const int et = threadIdx.x - 64;                      // 0..127
        // work items: (batch row b, unit pair up) -> item = b * UP + up ; thread handles items et, et + 128
        constexpr int UP = FWD_UNITS / 2;
        const int n_items = p.B * UP;
        float c[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
        int len_i[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int item = et + s * EPI_THREADS;
            len_i[s] = (item < n_items && p.lens) ? p.lens[item / UP] : p.T;
        }
        const int u0 = FWD_UNITS * cta;
        if (p.t0 > 0) {                                       // resume the cell state saved by the previous chunk
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int item = et + s * EPI_THREADS;
                if (item < n_items) {
                    const int b = item / UP, up = item % UP;
                    const float2 cv = *reinterpret_cast<const float2*>(p.cstate + (static_cast<long long>(p.t0 - 1) * p.B + b) * LH + u0 + 2 * up);
                    c[s][0] = cv.x; c[s][1] = cv.y;
                }
            }
        }
        float bias[2][8];                                     // kXIn: b_ih + b_hh of this thread's (gate, unit) pairs
        if (kXIn) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int item = et + s * EPI_THREADS;
#pragma unroll
                for (int j = 0; j < 8; ++j) bias[s][j] = 0.f;
                if (item < n_items) {
                    const int up = item % UP;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int r = g * LH + u0 + 2 * up;
                        bias[s][2 * g] = p.b_ih[r] + p.b_hh[r];
                        bias[s][2 * g + 1] = p.b_ih[r + 1] + p.b_hh[r + 1];
                    }
                }
            }
        }
        for (int t = p.t0; t < p.t1; ++t) {
            float x[2][8];                                    // [item][gate*2 + e]
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int item = et + s * EPI_THREADS;
                if (kXIn) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[s][j] = bias[s][j];
                } else if (item < n_items) {
                    const int b = item / UP, up = item % UP;
                    const float* src = p.xproj + (static_cast<long long>(t) * p.B + b) * LG + u0 + 2 * up;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float2 v = __ldg(reinterpret_cast<const float2*>(src + g * LH));
                        x[s][2 * g] = v.x; x[s][2 * g + 1] = v.y;
                    }
                }
            }
            if (t > 0) {
                if (q < FWD_N / 16) {                         // M = 64 accumulator: row r lives in TMEM lane 32 (r / 16) + r % 16
                    mbar_wait(accum_full, (t - tm0) & 1, p.status, 215);     // (measured on B200: tools/probe/m64_layout.cu)
                    tc_fence_after();
                    if (et == 64) FT_TRACE(p, t, 4);          // warp 4 == quadrant 0
                    const int r = 16 * q + lane;              // W-slice row = gate * UNITS + unit (lanes 16..31 hold nothing)
                    for (int h = 0; h < nq; ++h) {            // 32 batch columns at a time
                        float acc[32];
                        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 32, acc);
                        tmem_ld_wait();
                        if (lane < 16) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) sAcc[(h * 32 + j) * AP + r] = acc[j];   // [batch][row]: readers unchanged
                        }
                    }
                    tc_fence_before();
                }
                epi_bar();
                if (et == 0) FT_TRACE(p, t, 8);               // accumulator transposed into shared memory
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int item = et + s * EPI_THREADS;
                    if (item < n_items) {
                        const int b = item / UP, up = item % UP;
                        const float* a = sAcc + b * AP + 2 * up;
#pragma unroll
                        for (int g = 0; g < 4; ++g) { x[s][2 * g] += a[g * FWD_UNITS]; x[s][2 * g + 1] += a[g * FWD_UNITS + 1]; }
                    }
                }
            }
            float gi[2][2], gf[2][2], gg[2][2], go[2][2], hv[2][2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int item = et + s * EPI_THREADS;
                if (item < n_items) {
                    const int b = item / UP, up = item % UP;
                    const bool valid = t < len_i[s];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        gi[s][e] = sigmoid_f(x[s][0 + e]);
                        gf[s][e] = sigmoid_f(x[s][2 + e]);
                        gg[s][e] = tanh_f(x[s][4 + e]);
                        go[s][e] = sigmoid_f(x[s][6 + e]);
                        c[s][e] = gf[s][e] * c[s][e] + gi[s][e] * gg[s][e];
                        hv[s][e] = valid ? go[s][e] * tanh_f(c[s][e]) : 0.f;
                    }
                    const long long r = static_cast<long long>(t) * p.B + b;
                    const __half2 h2 = __floats2half2_rn(hv[s][0], hv[s][1]);
                    *reinterpret_cast<__half2*>(p.hseq + r * p.ldh + u0 + 2 * up) = h2;     // critical path: h_t first
                }
            }
            if (et == 0) FT_TRACE(p, t, 9);                   // cell computed, h_t stores issued
            epi_bar();                                        // all h_t stores of this CTA precede the release
            if (et == 0) {
                FT_TRACE(p, t, 10);
                red_release_add(&p.flags[(t - p.t0) * FWD_NCH + cta / FWD_PROD], 1);      // cumulative release (gpu scope)
                FT_TRACE(p, t, 7);
            }
            // off the critical path: tensors only the backward pass reads
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int item = et + s * EPI_THREADS;
                if (item < n_items) {
                    const int b = item / UP, up = item % UP;
                    const long long r = static_cast<long long>(t) * p.B + b;
                    if (p.h32) *reinterpret_cast<float2*>(p.h32 + r * p.ldh32 + u0 + 2 * up) = make_float2(hv[s][0], hv[s][1]);
                    if (p.gates) {
                        __half* gp = p.gates + r * LG + u0 + 2 * up;
                        *reinterpret_cast<__half2*>(gp) = __floats2half2_rn(gi[s][0], gi[s][1]);
                        *reinterpret_cast<__half2*>(gp + LH) = __floats2half2_rn(gf[s][0], gf[s][1]);
                        *reinterpret_cast<__half2*>(gp + 2 * LH) = __floats2half2_rn(gg[s][0], gg[s][1]);
                        *reinterpret_cast<__half2*>(gp + 3 * LH) = __floats2half2_rn(go[s][0], go[s][1]);
                    }
                    if (p.cstate) *reinterpret_cast<float2*>(p.cstate + r * LH + u0 + 2 * up) = make_float2(c[s][0], c[s][1]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<64>(tmem_base);
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

PFN_encodeTiled get_tma_encoder();     // gemm.cu

// 3-D view of a row-major [rows, 64*nchunks] 16-bit tensor as {64 elems, rows, chunks}; box {64, box_rows, box_chunks}.
// The shared-memory image of one box is chunk-major: [chunk][row][128 B], SWIZZLE_128B -- i.e. `box_chunks`
// consecutive K-major UMMA tiles.
static int make_tmap_chunks(CUtensorMap* out, const void* ptr, long long rows, int nchunks, long long ld, int box_rows,
                            int box_chunks) {
    PFN_encodeTiled enc = get_tma_encoder();
    if (!enc) return -1;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) return ft_set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch");
    cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(nchunks)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld * 2), 128};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_chunks)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ft_set_error("cuTensorMapEncodeTiled (3-D chunk view) failed");
    return 0;
}

static bool g_half_sm = false;       // 64-CTA forward kernel (for two concurrent half-batch launches)
void set_lstm_half_sm(int on) { g_half_sm = on != 0; }

// Steps [t0, t1) of one layer's forward recurrence on 1024 / UNITS CTAs.  `flags` needs (t1 - t0) * 16 ints.
// x16 != null: folded input projection (kXIn): x16 [T*B, kx] fp16 (row pitch ldx), wih16 [4096, kx] fp16, biases fp32.
template <int GS, int UNITS, bool kXIn>
static int launch_fwd_t(int T, int B, int t0, int t1, const float* xproj, const void* whh16, const int* lens, void* hseq16,
                        long long ldh, void* gates16, float* cstate, float* h32, long long ldh32, int* flags, cudaStream_t st,
                        const void* x16 = nullptr, long long ldx = 0, int kx = 0, const void* wih16 = nullptr,
                        const float* b_ih = nullptr, const float* b_hh = nullptr) {
    constexpr int NCHT = FWD_NCH + (kXIn ? XCH : 0);
    constexpr int N = 4 * UNITS, W_BYTES = NCHT * N * 128, CTAS = LH / UNITS;
    LstmFwdChunkParams p;
    p.T = T; p.B = B; p.Bbox = (B + 7) & ~7; p.t0 = t0; p.t1 = t1;
    p.xproj = xproj; p.b_ih = b_ih; p.b_hh = b_hh; p.lens = lens; p.hseq = static_cast<__half*>(hseq16); p.ldh = ldh;
    p.gates = static_cast<__half*>(gates16); p.cstate = cstate; p.h32 = h32; p.ldh32 = ldh32;
    p.flags = flags; p.status = ft_status_word(); p.trace = (t0 == 0 && t1 == T) ? g_lstm_trace : nullptr;
    const int slot = p.Bbox * 128, nq = (B + 31) / 32;
    const int smem = (FWD_NCH + (kXIn ? 2 * XCH : 0)) * slot + W_BYTES + nq * 32 * (N + 1) * 4 + 8 + 256 + 1024;
    if (smem > smem_optin()) return ft_set_error("lstm_fwd: not enough shared memory");
    CUtensorMap tmW, tmH, tmWx, tmX;
    if (make_tmap_2d(&tmW, whh16, FMT_F16, LG, LH, LH, KCH, UNITS)) return -1;
    if (make_tmap_chunks(&tmH, hseq16, static_cast<long long>(T) * B, FWD_NCH, ldh, p.Bbox, GS)) return -1;
    if (kXIn) {
        if (kx <= 0 || kx > XCH * KCH) return ft_set_error("lstm_fwd: folded input width must be in (0, 128]");
        if (make_tmap_2d(&tmWx, wih16, FMT_F16, LG, kx, kx, KCH, UNITS)) return -1;
        if (make_tmap_2d(&tmX, x16, FMT_F16, static_cast<long long>(T) * B, kx, ldx, KCH, p.Bbox)) return -1;
    } else {
        tmWx = tmW; tmX = tmH;                                   // unused
    }
    if (cudaMemsetAsync(flags, 0, sizeof(int) * (t1 - t0) * FWD_NCH, st) != cudaSuccess) return ft_set_error("lstm_fwd: memset failed");
    void* fn = reinterpret_cast<void*>(lstm_fwd_kernel<GS, UNITS, kXIn>);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    TimeScope ts("lstm_fwd", t1 - t0, B, 0, st);
    void* args[] = {&tmW, &tmH, &tmWx, &tmX, &p};
    cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(CTAS), dim3(LSTM_THREADS), args, smem, st);
    if (e != cudaSuccess) return ft_set_error(cudaGetErrorString(e));
    ft_count_launch(1);
    return ft_check_launch("lstm_fwd_kernel");
}

int launch_lstm_fwd(int T, int B, const float* xproj, const void* whh16, const int* lens, void* hseq16, long long ldh,
                    void* gates16, float* cstate, float* h32, long long ldh32, int* flags, cudaStream_t st) {
    if (T <= 0 || B <= 0) return 0;
    if (B > 64) return ft_set_error("lstm_fwd: batch > 64 per call not supported (split the batch)");
    if (g_half_sm && B <= 32) return launch_fwd_t<8, 16, false>(T, B, 0, T, xproj, whh16, lens, hseq16, ldh, gates16, cstate, h32, ldh32, flags, st);
    static int gs = -1;            // FT_LSTM_FWD_GS: K-chunks per TMA group (16 / gs groups per step), tuning knob
    if (gs < 0) { const char* e = getenv("FT_LSTM_FWD_GS"); gs = e ? atoi(e) : 8; }
    if (gs == 4) return launch_fwd_t<4, 8, false>(T, B, 0, T, xproj, whh16, lens, hseq16, ldh, gates16, cstate, h32, ldh32, flags, st);
    if (gs == 16) return launch_fwd_t<16, 8, false>(T, B, 0, T, xproj, whh16, lens, hseq16, ldh, gates16, cstate, h32, ldh32, flags, st);
    return launch_fwd_t<8, 8, false>(T, B, 0, T, xproj, whh16, lens, hseq16, ldh, gates16, cstate, h32, ldh32, flags, st);
}

// Whole sequence with the input projection folded in (narrow inputs, kx <= 128; x16 row 0..B-1 must be zero: step 0 skips the
// contraction).  128 CTAs (or 64 with the half-SM switch).
int launch_lstm_fwd_xin(int T, int B, const void* x16, long long ldx, int kx, const void* wih16, const float* b_ih, const float* b_hh,
                        const void* whh16, const int* lens, void* hseq16, long long ldh, void* gates16, float* cstate, float* h32,
                        long long ldh32, int* flags, cudaStream_t st) {
    if (T <= 0 || B <= 0) return 0;
    if (B > 64) return ft_set_error("lstm_fwd: batch > 64 per call not supported (split the batch)");
    if (!x16 || !wih16 || !b_ih || !b_hh) return ft_set_error("lstm_fwd_xin: NULL argument");
    if (g_half_sm && B <= 32)
        return launch_fwd_t<8, 16, true>(T, B, 0, T, nullptr, whh16, lens, hseq16, ldh, gates16, cstate, h32, ldh32, flags, st, x16, ldx, kx, wih16, b_ih, b_hh);
    return launch_fwd_t<8, 8, true>(T, B, 0, T, nullptr, whh16, lens, hseq16, ldh, gates16, cstate, h32, ldh32, flags, st, x16, ldx, kx, wih16, b_ih, b_hh);
}

// Steps [t0, t1) of a layer on 64 CTAs (layer pipeline, ar_step.cu).  `flags` needs (t1 - t0) * 16 ints.
int launch_lstm_fwd_chunk(int T, int B, int t0, int t1, const float* xproj, const void* whh16, const int* lens, void* hseq16,
                          long long ldh, void* gates16, float* cstate, int* flags, cudaStream_t st, float* h32, long long ldh32) {
    if (T <= 0 || B <= 0 || t1 <= t0) return 0;
    if (B > 64 || t0 < 0 || t1 > T) return ft_set_error("lstm_fwd_chunk: bad batch or step range");
    if (t0 > 0 && !cstate) return ft_set_error("lstm_fwd_chunk: resuming a chunk needs the saved cell states");
    return launch_fwd_t<8, 16, false>(T, B, t0, t1, xproj, whh16, lens, hseq16, ldh, gates16, cstate, h32, ldh32, flags, st);
}

// Steps [t0, t1) of a layer's BPTT with the split-K cluster kernel, B <= 32.  `flags` needs (t1 - t0) * 64 ints; `dc_carry`
// ([B,1024] fp32) is only touched when the range is a proper chunk.
static int launch_bwd4(int T, int B, int t0, int t1, const float* dh_ext, long long ldd, const void* whhT16, const void* gates16,
                       const float* cstate, const int* lens, void* dG16, float* dc_carry, int* flags, cudaStream_t st) {
    LstmBwdChunkParams p;
    p.T = T; p.B = B; p.Bbox = (B + 7) & ~7; p.gs = 8; p.ng = 2; p.nring = 0; p.t0 = t0; p.t1 = t1; p.dc_carry = dc_carry;
    const int slot = p.Bbox * 128;
    p.dh_ext = dh_ext; p.ldd = ldd; p.gates = static_cast<const __half*>(gates16); p.cstate = cstate; p.lens = lens;
    p.dG = static_cast<__half*>(dG16); p.flags = flags; p.status = ft_status_word();
    p.trace = (t0 == 0 && t1 == T) ? g_lstm_trace : nullptr;
    CUtensorMap tmWT, tmG;
    if (make_tmap_2d(&tmWT, whhT16, FMT_F16, LH, LG, LG, KCH, B4_UNITS)) return -1;
    if (make_tmap_chunks(&tmG, dG16, static_cast<long long>(T) * B, BWD_NCH, LG, p.Bbox, 8)) return -1;
    if (cudaMemsetAsync(flags, 0, sizeof(int) * (t1 - t0) * BWD_NCH, st) != cudaSuccess) return ft_set_error("lstm_bwd: memset failed");
    const int smem = B4_NCH * slot + B4_W_BYTES + B4_XCHG_FLOATS * 4 + 256 + 1024;
    cudaFuncSetAttribute(lstm_bwd4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B4_CTAS); cfg.blockDim = dim3(LSTM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = B4_CLUSTER; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeCooperative;
    attrs[1].val.cooperative = 1;
    static int coop = -1;          // FT_BWD_COOP=0: plain cluster launch (ncu cannot replay cooperative + cluster launches)
    if (coop < 0) { const char* e = getenv("FT_BWD_COOP"); coop = e ? atoi(e) : 1; }
    cfg.attrs = attrs; cfg.numAttrs = coop ? 2 : 1;
    TimeScope ts("lstm_bwd", t1 - t0, B, 0, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_bwd4_kernel, tmWT, tmG, p);
    if (e != cudaSuccess) {                                  // co-residency is implied by 64 CTAs <= SMs; retry without the attribute
        cudaGetLastError();
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, lstm_bwd4_kernel, tmWT, tmG, p);
    }
    if (e != cudaSuccess) return ft_set_error(cudaGetErrorString(e));
    ft_count_launch(1);
    return ft_check_launch("lstm_bwd4_kernel");
}

int launch_lstm_bwd_chunk(int T, int B, int t0, int t1, const float* dh_ext, long long ldd, const void* whhT16, const void* gates16,
                          const float* cstate, const int* lens, void* dG16, float* dc_carry, int* flags, cudaStream_t st) {
    if (T <= 0 || B <= 0 || t1 <= t0) return 0;
    if (B > 32 || t0 < 0 || t1 > T || !dc_carry) return ft_set_error("lstm_bwd_chunk: bad batch, step range or carry buffer");
    return launch_bwd4(T, B, t0, t1, dh_ext, ldd, whhT16, gates16, cstate, lens, dG16, dc_carry, flags, st);
}

int launch_lstm_bwd(int T, int B, const float* dh_ext, long long ldd, const void* whhT16, const void* gates16,
                    const float* cstate, const int* lens, void* dG16, int* flags, cudaStream_t st) {
    if (T <= 0 || B <= 0) return 0;
    if (B > 64) return ft_set_error("lstm_bwd: batch > 64 per call not supported (split the batch)");
    static int use_cluster = -1;
    if (use_cluster < 0) { const char* e = getenv("FT_LSTM_BWD_CLUSTER"); use_cluster = e ? atoi(e) : 1; }
    if (B <= 32 && use_cluster)
        return launch_bwd4(T, B, 0, T, dh_ext, ldd, whhT16, gates16, cstate, lens, dG16, nullptr, flags, st);
    LstmBwdParams p;
    p.T = T; p.B = B; p.Bbox = (B + 7) & ~7;
    const int slot = p.Bbox * 128;
    const int fixed = BWD_W_BYTES + 128 * 17 * 4 + 512 + 1024;
    const int avail = smem_optin() - fixed;
    int gs = 8;
    { const char* e = getenv("FT_LSTM_BWD_GS"); if (e) { int v = atoi(e); if (v == 2 || v == 4 || v == 8 || v == 16) gs = v; } }
    while (gs > 2 && 2 * gs * slot > avail) gs >>= 1;          // at least a 2-deep ring of groups (gs stays even)
    int nring = avail / (gs * slot);
    if (nring > 8) nring = 8;
    if (nring < 2) return ft_set_error("lstm_bwd: not enough shared memory for the dG ring");
    p.gs = gs; p.ng = BWD_NCH / gs; p.nring = nring;
    p.dh_ext = dh_ext; p.ldd = ldd; p.gates = static_cast<const __half*>(gates16); p.cstate = cstate; p.lens = lens;
    p.dG = static_cast<__half*>(dG16); p.flags = flags; p.status = ft_status_word(); p.trace = g_lstm_trace;
    CUtensorMap tmWT, tmG;
    if (make_tmap_2d(&tmWT, whhT16, FMT_F16, LH, LG, LG, KCH, BWD_UNITS)) return -1;
    if (make_tmap_chunks(&tmG, dG16, static_cast<long long>(T) * B, BWD_NCH, LG, p.Bbox, gs)) return -1;
    if (cudaMemsetAsync(flags, 0, sizeof(int) * T * BWD_NCH, st) != cudaSuccess) return ft_set_error("lstm_bwd: memset failed");
    const int smem = nring * gs * slot + fixed;
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
