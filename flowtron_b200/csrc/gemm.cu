// tcgen05 + TMA GEMM for sm_100a:  C[M,N] = act(A[M,K] * B[N,K]^T + bias (+bias2)) (+ C if beta)
//
//  * operands 16-bit (f16 / bf16, kind::f16, UMMA_K=16) or tf32 (fp32 storage, kind::tf32, UMMA_K=8),
//    fp32 accumulation in TMEM;
//  * operand tiles staged by TMA (SWIZZLE_128B) into a 3-stage mbarrier ring; 2 CTAs/SM so one CTA's
//    epilogue overlaps the other's main loop;
//  * each operand may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows], i.e. the
//    transposed view of a natural activation/gradient tensor) -- this is what lets dW = dY^T X and
//    dX = dY W run on the tensors exactly as the forward pass left them (no transposes);
//  * warp roles: w0 = TMA producer, w1 = TMEM alloc + single-thread MMA issue, w2..w5 = epilogue
//    (TMEM -> registers -> bias/tanh/accumulate -> 128-bit global stores, fp32 and/or 16-bit copy).
//
// Used for every non-recurrent contraction on the Flowtron hot path (LSTM input projections
// flowtron.py:654-655, attention Q/K/V projections :568-571, DenseLayer :461-464, 1x1 conv :768, and
// all their dgrad/wgrad counterparts).
#include "ptx.cuh"
#include "ft_internal.h"

namespace ft {

constexpr int BM = 128, BN = 128, STAGES = 3;
constexpr int TILE_BYTES = BM * 128;               // 128 rows x 128 B (one SW128 row per matrix row)
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM = STAGES * 2 * TILE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
    int M, N, K;
    int a_fmt, b_fmt;          // 0 f16, 1 bf16, 2 tf32
    int a_mn, b_mn;            // 1 = operand given MN-major ([K, rows] row-major)
    const float* bias;         // [N] or null
    const float* bias2;        // [N] or null
    int act;                   // 0 none, 1 tanh, 2 multiply by (1 - aux^2)  (tanh backward; aux = saved tanh output)
    const __half* aux16; long long ldaux;
    int beta;                  // 1: C32 += result (C32 read-modify-write)
    float alpha;               // scale on the accumulator before bias
    const float* alpha_ptr;    // optional device scalar multiplied into alpha (loss-scale undo)
    float* C32; long long ldc32;
    void* C16; long long ldc16; int c16_fmt;   // 0 f16, 1 bf16
    int* status;
};

template <bool kTf32>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * STAGES * TILE_BYTES);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* accum_full = bars + 2 * STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x, tile_m = blockIdx.y;
    constexpr int ELT = kTf32 ? 4 : 2;
    constexpr int BK = 128 / ELT;          // elements per 128-byte swizzle row
    constexpr int UK = 32 / ELT;           // K per tcgen05.mma
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Producer / MMA warps run their loops warp-uniformly and only the instruction issue is predicated on one elected
    // lane: operands of UTMALDG / UTCHMMA must live in uniform registers, and computing them inside an `if (lane == 0)`
    // region makes the compiler wrap every issue in an R2UR + ELECT "waterfall" loop (~60 clk per MMA, measured).
    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1, p.status, 101);
            uint8_t* a = sA + s * TILE_BYTES;
            uint8_t* b = sB + s * TILE_BYTES;
            if (elect_one()) {
                mbar_expect_tx(&full[s], 2 * TILE_BYTES);
                if (!p.a_mn) {
                    tma_load_2d(a, &tmA, &full[s], kb * BK, tile_m * BM);             // box {BK, 128 rows}
                } else {                                                                // box {BK mn-elems, BK k-rows} x (BM/BK)
                    for (int j = 0; j < BM / BK; ++j)
                        tma_load_2d(a + j * (BK * 128), &tmA, &full[s], tile_m * BM + j * BK, kb * BK);
                }
                if (!p.b_mn) {
                    tma_load_2d(b, &tmB, &full[s], kb * BK, tile_n * BN);
                } else {
                    for (int j = 0; j < BN / BK; ++j)
                        tma_load_2d(b + j * (BK * 128), &tmB, &full[s], tile_n * BN + j * BK, kb * BK);
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (one elected lane issues)
        const uint32_t idesc = umma_idesc(BM, BN, p.a_fmt, p.b_fmt, p.a_mn, p.b_mn);
        const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_base, 0);
        // K-major : rows 128 B apart, 8-row groups 1024 B apart (SBO); K step = 32 B inside the row.
        // MN-major: k-rows 128 B apart, 8-k-row groups 1024 B apart (SBO), MN atoms BK*128 B apart (LBO); K step = UK k-rows.
        const uint64_t da0 = p.a_mn ? umma_smem_desc(smem_u32(sA), BK * 128, 1024) : umma_smem_desc(smem_u32(sA), 16, 1024);
        const uint64_t db0 = p.b_mn ? umma_smem_desc(smem_u32(sB), BK * 128, 1024) : umma_smem_desc(smem_u32(sB), 16, 1024);
        const uint64_t ka = p.a_mn ? (UK * 128) >> 4 : 2, kbs = p.b_mn ? (UK * 128) >> 4 : 2;   // per-MMA K advance (16-byte units)
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES, ph = (kb / STAGES) & 1;
            mbar_wait(&full[s], ph, p.status, 102);
            tc_fence_after();
            const uint64_t da = da0 + static_cast<uint64_t>(s * (TILE_BYTES >> 4)), db = db0 + static_cast<uint64_t>(s * (TILE_BYTES >> 4));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BK / UK; ++k) {
                    const uint32_t acc = (kb | k) != 0;
                    if (kTf32) umma_tf32(tmem_d, da + k * ka, db + k * kbs, idesc, acc);
                    else       umma_f16(tmem_d, da + k * ka, db + k * kbs, idesc, acc);
                }
                umma_commit(&empty[s]);                 // frees the smem slot once these MMAs retire
                if (kb == num_kb - 1) umma_commit(accum_full);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------- epilogue warps (TMEM lane quadrant = warp % 4)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const long long m = static_cast<long long>(tile_m) * BM + row;
        mbar_wait(accum_full, 0, p.status, 103);
        tc_fence_after();
        const bool row_ok = m < p.M;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            const int n0 = tile_n * BN + c * 32;
            if (n0 >= p.N) break;                        // warp-uniform
            float v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
            tmem_ld_wait();
            if (!row_ok) continue;
            const int nvalid = min(32, p.N - n0);
            const float alpha = p.alpha_ptr ? p.alpha * __ldg(p.alpha_ptr) : p.alpha;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = v[j] * alpha;
                if (j < nvalid) {
                    if (p.bias) x += __ldg(p.bias + n0 + j);
                    if (p.bias2) x += __ldg(p.bias2 + n0 + j);
                }
                if (p.act == 1) x = tanh_f(x);
                v[j] = x;
            }
            if (p.act == 2) {
                const __half* ax = p.aux16 + m * p.ldaux + n0;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < nvalid) {
                        const float y = __half2float(ax[j]);
                        v[j] *= (1.f - y * y);
                    }
                }
            }
            if (p.C32) {
                float* dst = p.C32 + m * p.ldc32 + n0;
                if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        if (p.beta) { float4 c0 = d4[j]; o.x += c0.x; o.y += c0.y; o.z += c0.z; o.w += c0.w; }
                        d4[j] = o;
                        if (p.beta) { v[4 * j] = o.x; v[4 * j + 1] = o.y; v[4 * j + 2] = o.z; v[4 * j + 3] = o.w; }
                    }
                } else {
                    for (int j = 0; j < nvalid; ++j) {
                        float o = v[j];
                        if (p.beta) o += dst[j];
                        dst[j] = o;
                        v[j] = o;
                    }
                }
            }
            if (p.C16) {
                uint16_t* dst = reinterpret_cast<uint16_t*>(p.C16) + m * p.ldc16 + n0;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (p.c16_fmt == 0) {      // fp16: saturate instead of overflowing to inf
                        __half2 h = __floats2half2_rn(fminf(fmaxf(v[2 * j], -65504.f), 65504.f),
                                                      fminf(fmaxf(v[2 * j + 1], -65504.f), 65504.f));
                        pk[j] = *reinterpret_cast<uint32_t*>(&h);
                    } else {
                        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        pk[j] = *reinterpret_cast<uint32_t*>(&h);
                    }
                }
                if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                    for (int j = 0; j < 4; ++j) d4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                } else {
                    const uint16_t* ps = reinterpret_cast<const uint16_t*>(pk);
                    for (int j = 0; j < nvalid; ++j) dst[j] = ps[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<BN>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host
static PFN_encodeTiled g_encode = nullptr;

int ensure_tma_encoder() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return ft_set_error("cuTensorMapEncodeTiled entry point unavailable");
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

PFN_encodeTiled get_tma_encoder() { return ensure_tma_encoder() ? nullptr : g_encode; }

// 2D row-major tensor [rows, cols] (cols contiguous, row pitch ld elements), box {box_cols, box_rows}, SW128.
int make_tmap_2d(CUtensorMap* out, const void* ptr, int fmt, long long rows, long long cols, long long ld,
                 int box_cols, int box_rows) {
    if (ensure_tma_encoder()) return -1;
    const int elt = (fmt == 2) ? 4 : 2;
    CUtensorMapDataType dt = fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elt) & 15)) return ft_set_error("TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch");
    if (box_cols * elt != 128) return ft_set_error("TMA box inner extent must be 128 bytes (SWIZZLE_128B)");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld * elt)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ft_set_error("cuTensorMapEncodeTiled failed");
    return 0;
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
    const bool tf32 = (g.a_fmt == 2);
    if ((g.a_fmt == 2) != (g.b_fmt == 2)) return ft_set_error("gemm: tf32 must be used for both operands");
    if (tf32 && (g.a_mn || g.b_mn)) return ft_set_error("gemm: MN-major tf32 operands not supported");
    const int elt = tf32 ? 4 : 2, BK = 128 / elt;
    CUtensorMap tmA, tmB;
    int rc;
    if (!g.a_mn) rc = make_tmap_2d(&tmA, g.A, g.a_fmt, g.M, g.K, g.lda, BK, BM);
    else         rc = make_tmap_2d(&tmA, g.A, g.a_fmt, g.K, g.M, g.lda, BK, BK);
    if (rc) return rc;
    if (!g.b_mn) rc = make_tmap_2d(&tmB, g.B, g.b_fmt, g.N, g.K, g.ldb, BK, BN);
    else         rc = make_tmap_2d(&tmB, g.B, g.b_fmt, g.K, g.N, g.ldb, BK, BK);
    if (rc) return rc;
    GemmParams p;
    p.M = g.M; p.N = g.N; p.K = g.K; p.a_fmt = g.a_fmt; p.b_fmt = g.b_fmt; p.a_mn = g.a_mn; p.b_mn = g.b_mn;
    p.bias = g.bias; p.bias2 = g.bias2; p.act = g.act; p.alpha_ptr = g.alpha_ptr; p.aux16 = static_cast<const __half*>(g.aux16); p.ldaux = g.ldaux; p.beta = g.beta; p.alpha = g.alpha;
    p.C32 = g.C32; p.ldc32 = g.ldc32; p.C16 = g.C16; p.ldc16 = g.ldc16; p.c16_fmt = g.c16_fmt;
    p.status = ft_status_word();
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        cudaFuncSetAttribute(gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
        attr_set = true;
    }
    TimeScope ts(g.a_mn ? "gemm_wgrad" : (g.b_mn ? "gemm_dgrad" : "gemm_fwd"), g.M, g.N, g.K, st);
    if (tf32) gemm_kernel<true><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(tmA, tmB, p);
    else      gemm_kernel<false><<<grid, GEMM_THREADS, GEMM_SMEM, st>>>(tmA, tmB, p);
    ft_count_launch(1);
    return ft_check_launch("gemm_kernel");
}

}  // namespace ft
