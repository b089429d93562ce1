// Host-side orchestration of one Flowtron AR flow step (training direction) on a CUDA stream:
// AR_Step.forward / AR_Back_Step.forward (flowtron.py:725-773, 605-627) and their autograd.
// No allocation, no synchronisation: every intermediate lives in the caller's `saved` / `scratch` areas,
// whose layouts are planned by the same code that uses them (Plan below).
//
// Precision plan: forward tensor-core operands are fp16 (11-bit significand, same as tf32; every
// forward operand is either a weight or a bounded activation); backward operands are fp16 too, on gradients
// multiplied by a per-flow power-of-two loss scale chosen on the device (launch_grad_scale); accumulation, LSTM cell state, softmax, exp/log and all reductions are fp32.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ft_internal.h"
#include "../../include/flowtron_b200.h"

namespace ft {

namespace {

constexpr int H = 1024, G = 4096;

struct Region { const char* name; size_t off, bytes; };

struct Plan {
    uint8_t* base = nullptr;
    size_t off = 0;
    std::vector<Region>* regs = nullptr;
    template <typename T>
    T* get(const char* name, size_t n) {
        const size_t bytes = n * sizeof(T);
        const size_t o = off;
        off = (off + bytes + 255) & ~static_cast<size_t>(255);
        if (regs) regs->push_back({name, o, bytes});
        return base ? reinterpret_cast<T*>(base + o) : nullptr;
    }
};

struct Dims {
    long long R, RL;
    int T, B, L, M, A, E, D;
    explicit Dims(const FtArStepDesc& d)
        : R(static_cast<long long>(d.T) * d.B), RL(static_cast<long long>(d.L) * d.B), T(d.T), B(d.B), L(d.L), M(d.n_mel),
          A(d.n_attn), E(d.n_text), D(H + d.n_attn) {}
};

// ---- tensors that persist from forward to backward
struct Saved {
    uint16_t *mel_in16, *d16, *gatesA, *text16, *h0_16, *gates0, *h1_16, *gates1, *y1_16, *y2_16;
    float *mel_flow, *cA, *Kp, *Vp, *Q, *p_save, *c0, *c1, *o32;
    void plan(Plan& p, const FtArStepDesc& d) {
        const Dims n(d);
        mel_in16 = p.get<uint16_t>("mel_in16", n.R * n.M);
        mel_flow = d.reversed ? p.get<float>("mel_flow", n.R * n.M) : nullptr;
        d16 = p.get<uint16_t>("d16", n.R * n.D);
        gatesA = p.get<uint16_t>("gatesA", n.R * G);
        cA = p.get<float>("cA", n.R * H);
        text16 = p.get<uint16_t>("text16", n.RL * n.E);
        Kp = p.get<float>("Kp", n.RL * n.A);
        Vp = p.get<float>("Vp", n.RL * n.A);
        Q = p.get<float>("Q", n.R * n.A);
        p_save = d.has_prior ? p.get<float>("p_save", static_cast<size_t>(n.B) * n.T * n.L) : nullptr;
        h0_16 = p.get<uint16_t>("h0_16", n.R * H);
        gates0 = p.get<uint16_t>("gates0", n.R * G);
        c0 = p.get<float>("c0", n.R * H);
        h1_16 = p.get<uint16_t>("h1_16", n.R * H);
        gates1 = p.get<uint16_t>("gates1", n.R * G);
        c1 = p.get<float>("c1", n.R * H);
        y1_16 = p.get<uint16_t>("y1_16", n.R * H);
        y2_16 = p.get<uint16_t>("y2_16", n.R * H);
        o32 = p.get<float>("o32", n.R * 2 * n.M);
    }
};

// ---- 16-bit weight copies (fp16 for forward, bf16 for backward)
struct W16 {
    uint16_t *w_ih_a, *w_hh_a, *w_ih0, *w_hh0, *w_ih1, *w_hh1, *wq, *wk, *wv, *w1, *w2, *wc;
    void plan(Plan& p, const Dims& n) {
        w_ih_a = p.get<uint16_t>("w_ih_a", static_cast<size_t>(G) * n.M);
        w_hh_a = p.get<uint16_t>("w_hh_a", static_cast<size_t>(G) * H);
        w_ih0 = p.get<uint16_t>("w_ih0", static_cast<size_t>(G) * n.D);
        w_hh0 = p.get<uint16_t>("w_hh0", static_cast<size_t>(G) * H);
        w_ih1 = p.get<uint16_t>("w_ih1", static_cast<size_t>(G) * H);
        w_hh1 = p.get<uint16_t>("w_hh1", static_cast<size_t>(G) * H);
        wq = p.get<uint16_t>("wq", static_cast<size_t>(n.A) * H);
        wk = p.get<uint16_t>("wk", static_cast<size_t>(n.A) * n.E);
        wv = p.get<uint16_t>("wv", static_cast<size_t>(n.A) * n.E);
        w1 = p.get<uint16_t>("w1", static_cast<size_t>(H) * H);
        w2 = p.get<uint16_t>("w2", static_cast<size_t>(H) * H);
        wc = p.get<uint16_t>("wc", static_cast<size_t>(2 * n.M) * H);
    }
};

// Layer pipeline (default ON, FT_PIPE_FWD=0 / FT_PIPE_BWD=0 fall back to one launch per layer): lstm layers 0
// and 1 run as two 64-CTA recurrences one chunk of FT_PIPE_CHUNK steps apart, forward and BPTT.  Measured r2 on B200
// (B=32, T=1000): 81.97 -> 66.83 ms/step.
static bool pipe_fwd_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FT_PIPE_FWD"); v = (!e || atoi(e) != 0) ? 1 : 0; }
    return v == 1;
}
static bool pipe_bwd_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FT_PIPE_BWD"); v = (!e || atoi(e) != 0) ? 1 : 0; }
    return v == 1;
}
static int pipe_chunk_steps() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FT_PIPE_CHUNK"); v = e ? atoi(e) : 100; if (v < 8) v = 8; }
    return v;
}

struct FwdScratch {
    W16 w;
    float* X;
    int* flags;
    float* X1 = nullptr;        // second input-projection buffer (pipelined layers only)
    float *hA32 = nullptr, *ctx32 = nullptr;     // fp32 copies of [hA ; ctx] for the gate logits (gated flow only)
    void plan(Plan& p, const FtArStepDesc& d) {
        const Dims n(d);
        w.plan(p, n);
        X = p.get<float>("X", n.R * G);
        flags = p.get<int>("flags", static_cast<size_t>(n.T) * 64);
        if (pipe_fwd_enabled()) X1 = p.get<float>("X1", n.R * G);
        if (d.has_gate) {
            hA32 = p.get<float>("hA32", n.R * H);
            ctx32 = p.get<float>("ctx32", n.R * n.A);
        }
    }
};

// What the attention-LSTM part of the backward (ar_step_bwd_attn) needs from the main part (ar_step_bwd_main): the two are
// separate entry points so that the caller's autograd can run whatever depends only on d_text (the text encoder's backward)
// concurrently with the attention LSTM's BPTT, a 64-SM kernel.  Lives in caller memory (`carry`), not in the shared scratch.
struct BwdCarry {
    float *dd, *dmel_flow, *scale;
    void plan(Plan& p, const FtArStepDesc& d) {
        const Dims n(d);
        dd = p.get<float>("dd", n.R * n.D);                 // [R, D]: cols 0:H = dhA (attention-LSTM output gradient), H: = dctx
        dmel_flow = p.get<float>("dmel_flow", n.R * n.M);
        scale = p.get<float>("scale", 8);
    }
};

struct BwdScratch {
    W16 w;                      // fp16, natural layouts (w_hh* hold the TRANSPOSED recurrent weights [H, 4H])
    uint16_t *dG1, *dG0, *dGa, *do16, *dy2, *dy1, *dQ16, *dK16, *dV16;
    float *dh, *dQ, *dK, *dV, *dmel_in;
    int* flags;
    float *dcarry1 = nullptr, *dcarry0 = nullptr;      // dc*f hand-over between BPTT chunks (pipelined layers only)
    void plan(Plan& p, const FtArStepDesc& d) {
        const Dims n(d);
        w.plan(p, n);
        dG1 = p.get<uint16_t>("dG1", n.R * G);      // one per LSTM layer: the wgrad GEMMs of a layer run on the side
        dG0 = p.get<uint16_t>("dG0", n.R * G);      // stream while the next BPTT already writes its own dG
        dGa = p.get<uint16_t>("dGa", n.R * G);
        do16 = p.get<uint16_t>("do16", n.R * 2 * n.M);
        dy2 = p.get<uint16_t>("dy2", n.R * H);
        dy1 = p.get<uint16_t>("dy1", n.R * H);
        dQ16 = p.get<uint16_t>("dQ16", n.R * n.A);
        dK16 = p.get<uint16_t>("dK16", n.RL * n.A);
        dV16 = p.get<uint16_t>("dV16", n.RL * n.A);
        dh = p.get<float>("dh", n.R * H);
        dQ = p.get<float>("dQ", n.R * n.A);
        dK = p.get<float>("dK", n.RL * n.A);
        dV = p.get<float>("dV", n.RL * n.A);
        dmel_in = p.get<float>("dmel_in", n.R * n.M);
        flags = p.get<int>("flags", static_cast<size_t>(n.T) * 64);
        if (pipe_bwd_enabled()) {
            dcarry1 = p.get<float>("dcarry1", static_cast<size_t>(n.B) * H);
            dcarry0 = p.get<float>("dcarry0", static_cast<size_t>(n.B) * H);
        }
    }
};

int check_desc(const FtArStepDesc& d) {
    if (d.n_hidden != H) return ft_set_error("ar_step: this build supports n_hidden == 1024 only");
    if (d.T <= 0 || d.B <= 0 || d.L <= 0) return ft_set_error("ar_step: empty shape");
    if (d.B > 64) return ft_set_error("ar_step: batch > 64 per call not supported (split the batch)");
    if (d.n_mel % 8 || d.n_attn % 64 || d.n_text % 8) return ft_set_error("ar_step: n_mel %8, n_attn %64, n_text %8 required");
    if (d.L > 256) return ft_set_error("ar_step: L > 256 not supported");
    return 0;
}

#define FT_TRY(x) do { if ((x) != 0) return -1; } while (0)

// C = A[M,K] B[N,K]^T helpers -----------------------------------------------------------------------------
int gemm_fwd(cudaStream_t st, long long M, int N, int K, const void* A, long long lda, const void* B, long long ldb,
             const float* bias, const float* bias2, int act, float* C32, long long ldc32, void* C16, long long ldc16) {
    GemmArgs g;
    g.M = static_cast<int>(M); g.N = N; g.K = K;
    g.A = A; g.lda = lda; g.a_fmt = FMT_F16; g.B = B; g.ldb = ldb; g.b_fmt = FMT_F16;
    g.bias = bias; g.bias2 = bias2; g.act = act;
    g.C32 = C32; g.ldc32 = ldc32; g.C16 = C16; g.ldc16 = ldc16; g.c16_fmt = FMT_F16;
    return launch_gemm(g, st);
}
// dgrad: dX[M,K] = dY[M,N] W[N,K]      (A = dY K-major, B = W given MN-major)
int gemm_dgrad(cudaStream_t st, long long M, int Kout, int Nred, const void* dY, long long lddy, const void* W, long long ldw,
               int beta, float* C32, long long ldc32, void* C16, long long ldc16, const void* aux16, long long ldaux,
               const float* alpha_ptr = nullptr) {
    GemmArgs g;
    g.M = static_cast<int>(M); g.N = Kout; g.K = Nred;
    g.A = dY; g.lda = lddy; g.a_fmt = FMT_F16; g.B = W; g.ldb = ldw; g.b_fmt = FMT_F16; g.b_mn = 1;
    g.alpha_ptr = alpha_ptr;
    g.beta = beta; g.C32 = C32; g.ldc32 = ldc32; g.C16 = C16; g.ldc16 = ldc16; g.c16_fmt = FMT_F16;
    if (aux16) { g.act = 2; g.aux16 = aux16; g.ldaux = ldaux; }
    return launch_gemm(g, st);
}
// wgrad: dW[N,K] = dY[R,N]^T X[R,K]     (both operands MN-major views of the natural tensors)
int gemm_wgrad(cudaStream_t st, int N, int K, long long R, const void* dY, long long lddy, const void* X, long long ldx,
               float* dW, long long ldw, const float* inv_scale) {
    if (R <= 0) return cudaMemsetAsync(dW, 0, sizeof(float) * N * ldw, st) == cudaSuccess ? 0 : ft_set_error("memset failed");
    GemmArgs g;
    g.M = N; g.N = K; g.K = static_cast<int>(R);
    g.A = dY; g.lda = lddy; g.a_fmt = FMT_F16; g.a_mn = 1; g.B = X; g.ldb = ldx; g.b_fmt = FMT_F16; g.b_mn = 1;
    g.alpha_ptr = inv_scale;          // undo the loss scale on the way out
    g.C32 = dW; g.ldc32 = ldw;
    return launch_gemm(g, st);
}

// Side stream for work that is off the backward critical path (weight gradients, bias column sums): it runs under the
// next persistent BPTT kernel, which occupies 64 of the 148 SMs.  One side stream + event per caller stream; ordering is
// by events only (fork: side waits for "main so far"; join: main waits for "side so far").
struct Side {
    cudaStream_t s = nullptr;
    cudaEvent_t ev_main = nullptr, ev_side = nullptr;
};
static std::map<cudaStream_t, Side> g_sides;
static std::mutex g_side_mu;
static bool g_side_enabled = true;

static Side* get_side(cudaStream_t main) {
    if (!g_side_enabled) return nullptr;
    std::lock_guard<std::mutex> lk(g_side_mu);
    auto it = g_sides.find(main);
    if (it != g_sides.end()) return &it->second;
    Side sd;
    if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&sd.ev_main, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&sd.ev_side, cudaEventDisableTiming);
    g_sides[main] = sd;
    return &g_sides[main];
}
static cudaStream_t fork_side(Side* sd, cudaStream_t main) {      // returns the stream to launch off-path work on
    if (!sd) return main;
    cudaEventRecord(sd->ev_main, main);
    cudaStreamWaitEvent(sd->s, sd->ev_main, 0);
    return sd->s;
}
static void join_side(Side* sd, cudaStream_t main) {
    if (!sd) return;
    cudaEventRecord(sd->ev_side, sd->s);
    cudaStreamWaitEvent(main, sd->ev_side, 0);
}

// Streams of the experimental layer pipeline: sB runs the second recurrence, sC its per-chunk input-projection GEMMs.
struct Pipe {
    cudaStream_t sA = nullptr, sB = nullptr, sC = nullptr, sD = nullptr;      // sA / sB: HIGH-priority streams of the two recurrences, sC: their GEMMs,
                                                                              // sD: lowest priority, work nobody waits for yet
    cudaEvent_t evA = nullptr, evG = nullptr, evB = nullptr, ev0 = nullptr, evJ = nullptr, evD = nullptr;
};
// FT_PIPE_PRIO (default 1): the chunk kernels of the pipelined recurrences are launched on high-priority streams, so that when a
// chunk ends its 64 SMs go to the next chunk (a cooperative launch that needs all of them at once) and not to queued CTAs of the
// GEMM / attention kernels that share the machine.
static bool pipe_prio() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FT_PIPE_PRIO"); v = (!e || atoi(e) != 0) ? 1 : 0; }
    return v == 1;
}
static std::map<cudaStream_t, Pipe> g_pipes;
static Pipe* get_pipe(cudaStream_t main) {
    std::lock_guard<std::mutex> lk(g_side_mu);
    auto it = g_pipes.find(main);
    if (it != g_pipes.end()) return &it->second;
    Pipe pp;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);            // hi is numerically smaller
    const int pr = pipe_prio() ? prio_hi : prio_lo;
    if (cudaStreamCreateWithPriority(&pp.sA, cudaStreamNonBlocking, pr) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&pp.sB, cudaStreamNonBlocking, pr) != cudaSuccess) return nullptr;
    const int mid = pipe_prio() && prio_hi < prio_lo - 1 ? prio_lo - 1 : prio_lo;      // the pipeline's own GEMMs beat sD's
    if (cudaStreamCreateWithPriority(&pp.sC, cudaStreamNonBlocking, mid) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&pp.sD, cudaStreamNonBlocking, prio_lo) != cudaSuccess) return nullptr;
    cudaEventCreateWithFlags(&pp.evJ, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pp.evD, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pp.evA, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pp.evG, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pp.evB, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pp.ev0, cudaEventDisableTiming);
    g_pipes[main] = pp;
    return &g_pipes[main];
}

int zero(void* p, size_t bytes, cudaStream_t st) {
    return cudaMemsetAsync(p, 0, bytes, st) == cudaSuccess ? 0 : ft_set_error("cudaMemsetAsync failed");
}
int copy_f32(float* dst, const float* src, size_t n, cudaStream_t st) {
    return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st) == cudaSuccess ? 0 : ft_set_error("cudaMemcpyAsync failed");
}

}  // namespace

// =================================================================================================== forward
static thread_local cudaEvent_t g_text_ready = nullptr;
void set_text_ready_event(void* ev) { g_text_ready = static_cast<cudaEvent_t>(ev); }

int ar_step_fwd(const FtArStepDesc& d, const FtArStepWeights& w, const float* mel, const float* text, const int* in_lens,
                const int* out_lens, const float* prior, float* mel_out, float* log_s, float* gates, float* attn,
                float* logprob, void* saved, void* scratch, cudaStream_t st) {
    cudaEvent_t text_ready = g_text_ready;        // one-shot: consumed by this call whatever happens next
    g_text_ready = nullptr;
    FT_TRY(check_desc(d));
    if (d.has_prior && !prior) return ft_set_error("ar_step_fwd: has_prior set but attn_prior is NULL");
    if (d.has_gate && (!gates || !w.gate_w)) return ft_set_error("ar_step_fwd: has_gate set but gate pointers are NULL");
    const Dims n(d);
    Plan ps; ps.base = static_cast<uint8_t*>(saved);
    Saved S; S.plan(ps, d);
    Plan pf; pf.base = static_cast<uint8_t*>(scratch);
    FwdScratch F; F.plan(pf, d);

    // fp16 operand copies of the weights (61 M params -> ~0.1 ms; redone every call so optimizer updates are seen)
    FT_TRY(launch_cast(w.attn_lstm_w_ih, 2, F.w.w_ih_a, 0, static_cast<long long>(G) * n.M, st));
    FT_TRY(launch_cast(w.attn_lstm_w_hh, 2, F.w.w_hh_a, 0, static_cast<long long>(G) * H, st));
    FT_TRY(launch_cast(w.lstm_w_ih0, 2, F.w.w_ih0, 0, static_cast<long long>(G) * n.D, st));
    FT_TRY(launch_cast(w.lstm_w_hh0, 2, F.w.w_hh0, 0, static_cast<long long>(G) * H, st));
    FT_TRY(launch_cast(w.lstm_w_ih1, 2, F.w.w_ih1, 0, static_cast<long long>(G) * H, st));
    FT_TRY(launch_cast(w.lstm_w_hh1, 2, F.w.w_hh1, 0, static_cast<long long>(G) * H, st));
    FT_TRY(launch_cast(w.att_query, 2, F.w.wq, 0, static_cast<long long>(n.A) * H, st));
    FT_TRY(launch_cast(w.att_key, 2, F.w.wk, 0, static_cast<long long>(n.A) * n.E, st));
    FT_TRY(launch_cast(w.att_value, 2, F.w.wv, 0, static_cast<long long>(n.A) * n.E, st));
    FT_TRY(launch_cast(w.dense_w0, 2, F.w.w1, 0, static_cast<long long>(H) * H, st));
    FT_TRY(launch_cast(w.dense_w1, 2, F.w.w2, 0, static_cast<long long>(H) * H, st));
    FT_TRY(launch_cast(w.conv_w, 2, F.w.wc, 0, static_cast<long long>(2 * n.M) * H, st));

    // teacher-forcing shift (+ AR_Back_Step time reversal folded into the gather)
    FT_TRY(launch_prep_mel(mel, out_lens, n.T, n.B, n.M, d.reversed, S.mel_in16, S.mel_flow, st));
    const float* mel_flow = d.reversed ? S.mel_flow : mel;

    // Attention-LSTM / attention overlap (FT_ATT_OVERLAP = chunk length in steps, a multiple of 64; default 256, 0 disables; B <= 32 and
    // T >= chunk + 64): the attention LSTM runs as a 64-CTA recurrence in chunks; as soon as a chunk's h is there, a second stream
    // runs that chunk's query projection, fused attention, gate logits and lstm layer 0's input projection on the 84 free SMs,
    // underneath the next chunk's recurrence.  Only the last chunk's share of that work stays exposed (attention 0.95 ms + GEMMs
    // 0.45 ms per flow were serial).  The 64-CTA kernel has no room for the folded 80-channel projection next to its 128 KB weight
    // slice, so the projection is a per-chunk GEMM on a third stream, one chunk ahead.  Measured on B200, B=32, T=1000
    // (r2 calls 12 / 30): chunks of 128 LOSE (53.1 vs 51.1 ms/step before the GEMM epilogue rewrite, 49.1 vs 48.6 after: every chunk
    // boundary costs a kernel drain + launch + 128 KB weight reload on 64 SMs), chunks of 192 / 256 / 320 / 384 / 512 give
    // 47.9 / 46.4 / 47.2 / 46.4 / 46.5 against 48.6 serial.
    static int att_chunk = -1;
    if (att_chunk < 0) { const char* e = getenv("FT_ATT_OVERLAP"); att_chunk = e ? atoi(e) : 256; if (att_chunk % 64) att_chunk = 0; }
    Pipe* pa = (att_chunk > 0 && F.X1 && n.B <= 32 && n.T >= att_chunk + 64) ? get_pipe(st) : nullptr;
    bool xproj0_done = false;
    auto attention_rows = [&](cudaStream_t s2, int t0, int t1) -> int {      // Q projection, attention, gate logits of steps [t0, t1)
        const long long r0 = static_cast<long long>(t0) * n.B, rows = static_cast<long long>(t1 - t0) * n.B;
        FT_TRY(gemm_fwd(s2, rows, n.A, H, S.d16 + r0 * n.D, n.D, F.w.wq, H, nullptr, nullptr, 0, S.Q + r0 * n.A, n.A, nullptr, 0));
        AttnFwdArgs a;
        a.T = n.T; a.B = n.B; a.L = n.L; a.A = n.A;
        a.Q = S.Q; a.ldq = n.A; a.K = S.Kp; a.ldk = n.A; a.V = S.Vp; a.ldv = n.A; a.v = w.att_v;
        a.in_lens = in_lens; a.out_lens = out_lens; a.prior = d.has_prior ? prior : nullptr; a.reversed = d.reversed;
        a.temperature = d.temperature; a.attn = attn; a.logprob = logprob; a.p_save = S.p_save;
        a.ctx16 = S.d16 + H; a.ldc = n.D; a.ctx32 = F.ctx32; a.ldc32 = n.A;
        a.t_begin = t0; a.t_end = t1;
        FT_TRY(launch_attn_fwd(a, s2));
        if (d.has_gate) FT_TRY(launch_gate_fwd_f32(F.hA32 + r0 * H, H, H, F.ctx32 + r0 * n.A, n.A, n.A, w.gate_w, w.gate_b, rows, gates + r0, s2));
        return 0;
    };
    auto kv_projections = [&](cudaStream_t s2) -> int {
        // First use of `text`: if the caller produced it on another stream it handed us the event to wait for, so the encoder
        // ran underneath everything above (ft_ar_step_set_text_ready_event).
        if (text_ready) {
            if (cudaStreamWaitEvent(s2, text_ready, 0) != cudaSuccess) return ft_set_error("ar_step_fwd: cudaStreamWaitEvent(text_ready) failed");
        }
        FT_TRY(launch_cast(text, 2, S.text16, 0, n.RL * n.E, s2));
        FT_TRY(gemm_fwd(s2, n.RL, n.A, n.E, S.text16, n.E, F.w.wk, n.E, nullptr, nullptr, 0, S.Kp, n.A, nullptr, 0));
        FT_TRY(gemm_fwd(s2, n.RL, n.A, n.E, S.text16, n.E, F.w.wv, n.E, nullptr, nullptr, 0, S.Vp, n.A, nullptr, 0));
        return 0;
    };
    if (pa) {
        const int CH = att_chunk;
        auto xa = [&](cudaStream_t s2, int t0, int t1) {              // attention-LSTM input projection of steps [t0, t1) -> F.X1 rows
            const long long r0 = static_cast<long long>(t0) * n.B, rows = static_cast<long long>(t1 - t0) * n.B;
            return gemm_fwd(s2, rows, G, n.M, S.mel_in16 + r0 * n.M, n.M, F.w.w_ih_a, n.M, w.attn_lstm_b_ih, w.attn_lstm_b_hh, 0,
                            F.X1 + r0 * G, G, nullptr, 0);
        };
        cudaStream_t sr = pipe_prio() ? pa->sA : st;                 // the attention LSTM's stream
        cudaEventRecord(pa->ev0, st);
        if (sr != st) cudaStreamWaitEvent(sr, pa->ev0, 0);
        cudaStreamWaitEvent(pa->sB, pa->ev0, 0);
        cudaStreamWaitEvent(pa->sC, pa->ev0, 0);
        FT_TRY(kv_projections(pa->sC));
        FT_TRY(xa(sr, 0, CH < n.T ? CH : n.T));
        for (int t0 = 0; t0 < n.T; t0 += CH) {
            const int t1 = t0 + CH < n.T ? t0 + CH : n.T;
            if (t1 < n.T) {                                            // next chunk's projection, one chunk ahead, on sB
                FT_TRY(xa(pa->sB, t1, t1 + CH < n.T ? t1 + CH : n.T));
                cudaEventRecord(pa->evB, pa->sB);
            }
            FT_TRY(launch_lstm_fwd_chunk(n.T, n.B, t0, t1, F.X1, F.w.w_hh_a, out_lens, S.d16, n.D, S.gatesA, S.cA, F.flags, sr, F.hA32, H));
            cudaEventRecord(pa->evA, sr);
            if (t1 < n.T) cudaStreamWaitEvent(sr, pa->evB, 0);        // the next launch on sr needs its projection
            cudaStreamWaitEvent(pa->sC, pa->evA, 0);
            FT_TRY(attention_rows(pa->sC, t0, t1));
            const long long r0 = static_cast<long long>(t0) * n.B, rows = static_cast<long long>(t1 - t0) * n.B;
            FT_TRY(gemm_fwd(pa->sC, rows, G, n.D, S.d16 + r0 * n.D, n.D, F.w.w_ih0, n.D, w.lstm_b_ih0, w.lstm_b_hh0, 0, F.X + r0 * G, G, nullptr, 0));
        }
        cudaEventRecord(pa->evG, pa->sC);
        cudaStreamWaitEvent(st, pa->evG, 0);
        if (sr != st) { cudaEventRecord(pa->evJ, sr); cudaStreamWaitEvent(st, pa->evJ, 0); }
        xproj0_done = true;
    } else {
    // attention_lstm: the 80-channel input projection is folded into the persistent recurrence (no [R,4096] projection
    // tensor: -0.5 ms GEMM, -1 GB of HBM traffic per flow); h lands in d16[:, 0:H].  FT_LSTM_XIN=0: separate GEMM (r1 path).
    static int xin = -1;
    if (xin < 0) { const char* e = getenv("FT_LSTM_XIN"); xin = (!e || atoi(e) != 0) ? 1 : 0; }
    if (xin && n.M <= 128 && n.B <= 32) {      // B > 32: the double-buffered x tile no longer fits next to a 64-row h tile
        FT_TRY(launch_lstm_fwd_xin(n.T, n.B, S.mel_in16, n.M, n.M, F.w.w_ih_a, w.attn_lstm_b_ih, w.attn_lstm_b_hh, F.w.w_hh_a, out_lens,
                                   S.d16, n.D, S.gatesA, S.cA, F.hA32, H, F.flags, st));
    } else {
        FT_TRY(gemm_fwd(st, n.R, G, n.M, S.mel_in16, n.M, F.w.w_ih_a, n.M, w.attn_lstm_b_ih, w.attn_lstm_b_hh, 0, F.X, G, nullptr, 0));
        FT_TRY(launch_lstm_fwd(n.T, n.B, F.X, F.w.w_hh_a, out_lens, S.d16, n.D, S.gatesA, S.cA, F.hA32, H, F.flags, st));
    }
    // attention: K/V/Q projections, fused score+softmax(+prior)+context; ctx lands in d16[:, H:H+A]
    FT_TRY(kv_projections(st));
    FT_TRY(attention_rows(st, 0, n.T));
    }

    // 2-layer lstm
    Pipe* pp = (F.X1 && n.B <= 32 && n.T > pipe_chunk_steps()) ? get_pipe(st) : nullptr;
    if (pp) {
        // EXPERIMENTAL: layer 1 at step t needs layer 0 only up to t, so the two recurrences run as two 64-CTA kernels
        // one chunk apart: st = layer 0, sC = layer-1 input projection of a finished chunk, sB = layer 1.
        const int S_ = pipe_chunk_steps();
        int* flagsA = F.flags;
        int* flagsB = F.flags + static_cast<size_t>(S_) * 16;
        if (!xproj0_done) FT_TRY(gemm_fwd(st, n.R, G, n.D, S.d16, n.D, F.w.w_ih0, n.D, w.lstm_b_ih0, w.lstm_b_hh0, 0, F.X, G, nullptr, 0));
        cudaStream_t sr = pipe_prio() ? pp->sA : st;                 // layer 0's recurrence stream
        cudaEventRecord(pp->ev0, st);
        if (sr != st) cudaStreamWaitEvent(sr, pp->ev0, 0);
        cudaStreamWaitEvent(pp->sB, pp->ev0, 0);
        cudaStreamWaitEvent(pp->sC, pp->ev0, 0);
        for (int t0 = 0; t0 < n.T; t0 += S_) {
            const int t1 = t0 + S_ < n.T ? t0 + S_ : n.T;
            const long long r0 = static_cast<long long>(t0) * n.B, rows = static_cast<long long>(t1 - t0) * n.B;
            FT_TRY(launch_lstm_fwd_chunk(n.T, n.B, t0, t1, F.X, F.w.w_hh0, out_lens, S.h0_16, H, S.gates0, S.c0, flagsA, sr));
            cudaEventRecord(pp->evA, sr);
            cudaStreamWaitEvent(pp->sC, pp->evA, 0);
            FT_TRY(gemm_fwd(pp->sC, rows, G, H, S.h0_16 + r0 * H, H, F.w.w_ih1, H, w.lstm_b_ih1, w.lstm_b_hh1, 0,
                            F.X1 + r0 * G, G, nullptr, 0));
            cudaEventRecord(pp->evG, pp->sC);
            cudaStreamWaitEvent(pp->sB, pp->evG, 0);
            FT_TRY(launch_lstm_fwd_chunk(n.T, n.B, t0, t1, F.X1, F.w.w_hh1, out_lens, S.h1_16, H, S.gates1, S.c1, flagsB, pp->sB));
        }
        cudaEventRecord(pp->evB, pp->sB);
        cudaStreamWaitEvent(st, pp->evB, 0);
        if (sr != st) { cudaEventRecord(pp->evJ, sr); cudaStreamWaitEvent(st, pp->evJ, 0); }
    } else {
        if (!xproj0_done) FT_TRY(gemm_fwd(st, n.R, G, n.D, S.d16, n.D, F.w.w_ih0, n.D, w.lstm_b_ih0, w.lstm_b_hh0, 0, F.X, G, nullptr, 0));
        FT_TRY(launch_lstm_fwd(n.T, n.B, F.X, F.w.w_hh0, out_lens, S.h0_16, H, S.gates0, S.c0, nullptr, 0, F.flags, st));
        FT_TRY(gemm_fwd(st, n.R, G, H, S.h0_16, H, F.w.w_ih1, H, w.lstm_b_ih1, w.lstm_b_hh1, 0, F.X, G, nullptr, 0));
        FT_TRY(launch_lstm_fwd(n.T, n.B, F.X, F.w.w_hh1, out_lens, S.h1_16, H, S.gates1, S.c1, nullptr, 0, F.flags, st));
    }

    // dense x2 (tanh fused in the GEMM epilogue), 1x1 conv, affine coupling
    FT_TRY(gemm_fwd(st, n.R, H, H, S.h1_16, H, F.w.w1, H, w.dense_b0, nullptr, 1, nullptr, 0, S.y1_16, H));
    FT_TRY(gemm_fwd(st, n.R, H, H, S.y1_16, H, F.w.w2, H, w.dense_b1, nullptr, 1, nullptr, 0, S.y2_16, H));
    FT_TRY(gemm_fwd(st, n.R, 2 * n.M, H, S.y2_16, H, F.w.wc, H, w.conv_b, nullptr, 0, S.o32, 2 * n.M, nullptr, 0));
    FT_TRY(launch_affine_fwd(S.o32, mel_flow, out_lens, n.T, n.B, n.M, d.reversed, mel_out, log_s, st));
    return 0;
}

// =================================================================================================== backward
// Main part: everything except the attention LSTM.  Outputs d_text and every parameter gradient but attention_lstm's; leaves
// dhA / dmel_flow / the loss scale in `carry` for ar_step_bwd_attn.
int ar_step_bwd_main(const FtArStepDesc& d, const FtArStepWeights& w, const float* mel, const int* in_lens, const int* out_lens,
                     const float* attn, const float* d_mel_out, const float* d_log_s, const float* d_gates, const float* d_attn,
                     const float* d_logprob, float* d_text, const FtArStepWeights& g, void* saved, void* scratch, void* carry,
                     cudaStream_t st, bool may_fuse_bptt = false, bool* bptt_done = nullptr) {
    if (bptt_done) *bptt_done = false;
    FT_TRY(check_desc(d));
    const Dims n(d);
    Plan ps; ps.base = static_cast<uint8_t*>(saved);
    Saved S_; S_.plan(ps, d);
    Plan pb; pb.base = static_cast<uint8_t*>(scratch);
    BwdScratch F; F.plan(pb, d);
    Plan pc; pc.base = static_cast<uint8_t*>(carry);
    BwdCarry C; C.plan(pc, d);
    const float* mel_flow = d.reversed ? S_.mel_flow : mel;
    const long long R = n.R, RL = n.RL, Rm = n.R - n.B;     // Rm: rows with a predecessor step

    // fp16 operand copies: natural layouts for dgrad (B operand MN-major), transposed recurrent weights for BPTT
    FT_TRY(launch_cast(w.lstm_w_ih0, 2, F.w.w_ih0, 0, static_cast<long long>(G) * n.D, st));
    FT_TRY(launch_cast(w.lstm_w_ih1, 2, F.w.w_ih1, 0, static_cast<long long>(G) * H, st));
    FT_TRY(launch_cast(w.att_query, 2, F.w.wq, 0, static_cast<long long>(n.A) * H, st));
    FT_TRY(launch_cast(w.att_key, 2, F.w.wk, 0, static_cast<long long>(n.A) * n.E, st));
    FT_TRY(launch_cast(w.att_value, 2, F.w.wv, 0, static_cast<long long>(n.A) * n.E, st));
    FT_TRY(launch_cast(w.dense_w0, 2, F.w.w1, 0, static_cast<long long>(H) * H, st));
    FT_TRY(launch_cast(w.dense_w1, 2, F.w.w2, 0, static_cast<long long>(H) * H, st));
    FT_TRY(launch_cast(w.conv_w, 2, F.w.wc, 0, static_cast<long long>(2 * n.M) * H, st));
    FT_TRY(launch_transpose_cast_f16(w.lstm_w_hh0, F.w.w_hh0, G, H, st));
    FT_TRY(launch_transpose_cast_f16(w.lstm_w_hh1, F.w.w_hh1, G, H, st));

    // 0. loss scale for this flow's backward: S = 2^k with S * max|incoming grad| ~ 64 (device-side, no host sync)
    const long long RM = R * n.M;
    const long long BTL = static_cast<long long>(n.B) * n.T * n.L;
    FT_TRY(launch_grad_scale(d_mel_out, d_mel_out ? RM : 0, d_log_s, d_log_s ? RM : 0, d_gates, d_gates ? R : 0, 64.f, C.scale, st,
                             d_attn, d_attn ? BTL : 0, d_logprob, d_logprob ? BTL : 0));
    const float* S = C.scale;            // S[0] = scale, S[1] = 1/scale
    const float* iS = C.scale + 1;

    static int side_env = -1;
    if (side_env < 0) { const char* e = getenv("FT_SIDE_STREAM"); side_env = e ? atoi(e) : 1; g_side_enabled = side_env != 0; }
    Side* sd = get_side(st);
    cudaStream_t ss;

    // 1. affine coupling (scales the incoming gradients by S)
    FT_TRY(launch_affine_bwd(d_mel_out, d_log_s, S_.o32, mel_flow, out_lens, n.T, n.B, n.M, d.reversed, F.do16, C.dmel_flow, S, st));

    // 2. 1x1 conv     (weight / bias gradients go to the side stream, the dgrad chain stays on `st`)
    ss = fork_side(sd, st);
    FT_TRY(gemm_wgrad(ss, 2 * n.M, H, R, F.do16, 2 * n.M, S_.y2_16, H, g.conv_w, H, iS));
    FT_TRY(launch_colsum(F.do16, 0, 2 * n.M, R, 2 * n.M, g.conv_b, iS, ss));
    FT_TRY(gemm_dgrad(st, R, H, 2 * n.M, F.do16, 2 * n.M, F.w.wc, H, 0, nullptr, 0, F.dy2, H, S_.y2_16, H));     // * (1 - y2^2)

    // 3. dense layer 1 (second linear)
    ss = fork_side(sd, st);
    FT_TRY(gemm_wgrad(ss, H, H, R, F.dy2, H, S_.y1_16, H, g.dense_w1, H, iS));
    FT_TRY(launch_colsum(F.dy2, 0, H, R, H, g.dense_b1, iS, ss));
    FT_TRY(gemm_dgrad(st, R, H, H, F.dy2, H, F.w.w2, H, 0, nullptr, 0, F.dy1, H, S_.y1_16, H));                   // * (1 - y1^2)

    // 4. dense layer 0
    ss = fork_side(sd, st);
    FT_TRY(gemm_wgrad(ss, H, H, R, F.dy1, H, S_.h1_16, H, g.dense_w0, H, iS));
    FT_TRY(launch_colsum(F.dy1, 0, H, R, H, g.dense_b0, iS, ss));
    FT_TRY(gemm_dgrad(st, R, H, H, F.dy1, H, F.w.w1, H, 0, F.dh, H, nullptr, 0, nullptr, 0));

    FT_TRY(zero(F.dK, sizeof(float) * RL * n.A, st));
    FT_TRY(zero(F.dV, sizeof(float) * RL * n.A, st));
    FT_TRY(zero(g.att_v, sizeof(float) * n.A, st));
    if (d.has_gate && g.gate_w) {
        FT_TRY(zero(g.gate_w, sizeof(float) * n.D, st));
        FT_TRY(zero(g.gate_b, sizeof(float), st));
    }
    // Steps 6b-9a for the time steps [t0, t1): lstm layer 0's input gradient (dd = [dhA ; dctx]), the gate layer, the attention
    // backward (tanh recompute; dK / dV / dv accumulate over calls) and the query projection's share of dhA.
    auto attention_bwd_rows = [&](cudaStream_t s2, int t0, int t1) -> int {
        const long long r0 = static_cast<long long>(t0) * n.B, rows = static_cast<long long>(t1 - t0) * n.B;
        FT_TRY(gemm_dgrad(s2, rows, n.D, G, F.dG0 + r0 * G, G, F.w.w_ih0, n.D, 0, C.dd + r0 * n.D, n.D, nullptr, 0, nullptr, 0));
        // 7. gate layer (last flow only): dd is in the scaled domain, the gate's own parameter gradients are not
        if (d.has_gate && g.gate_w && d_gates)
            FT_TRY(launch_gate_bwd(S_.d16 + r0 * n.D, n.D, n.D, w.gate_w, d_gates + r0, rows, C.dd + r0 * n.D, n.D, g.gate_w, g.gate_b, S, s2));
        // 8. attention (score/softmax/context) with tanh recompute
        AttnBwdArgs a;
        a.T = n.T; a.B = n.B; a.L = n.L; a.A = n.A;
        a.Q = S_.Q; a.ldq = n.A; a.K = S_.Kp; a.ldk = n.A; a.V = S_.Vp; a.ldv = n.A; a.v = w.att_v;
        a.in_lens = in_lens; a.out_lens = out_lens; a.attn = attn; a.p_save = S_.p_save; a.temperature = d.temperature;
        a.dctx = C.dd + H; a.lddc = n.D; a.dattn_ext = d_attn; a.dlp_ext = d_logprob; a.scale = S;
        a.dQ = F.dQ; a.lddq = n.A; a.dK = F.dK; a.lddk = n.A; a.dV = F.dV; a.lddv = n.A; a.dv = g.att_v;
        a.t_begin = t0; a.t_end = t1;
        FT_TRY(launch_attn_bwd(a, s2));
        // 9a. query projection (d16 = [hA ; ctx] saved in fp16, row pitch D)
        FT_TRY(launch_cast(F.dQ + r0 * n.A, 2, F.dQ16 + r0 * n.A, 0, rows * n.A, s2));
        FT_TRY(gemm_dgrad(s2, rows, H, n.A, F.dQ16 + r0 * n.A, n.A, F.w.wq, H, 1, C.dd + r0 * n.D, n.D, nullptr, 0, nullptr, 0));   // dhA += dQ Wq
        return 0;
    };
    // FT_ATTB_PIPE=1 (needs FT_PIPE_CHUNK % 64 == 0): the attention backward of every chunk lstm layer 0's BPTT has finished runs on a
    // lowest-priority stream on whatever the two recurrences and their GEMMs leave free, instead of after the pipeline.
    static int attb_pipe = -1;
    if (attb_pipe < 0) { const char* e = getenv("FT_ATTB_PIPE"); attb_pipe = e ? atoi(e) : 0; }
    bool attention_bwd_done = false;
    Pipe* pp = (F.dcarry1 && n.B <= 32 && 2 * pipe_chunk_steps() <= n.T) ? get_pipe(st) : nullptr;
    if (pp) {
        // EXPERIMENTAL 5+6: BPTT of layer 1 (st) and layer 0 (sB) one chunk apart; the layer-1 dgrad of a finished chunk
        // (sC) turns dG1 rows into layer 0's incoming dh rows (F.dh is reused row-disjointly: layer 1 only reads rows
        // below the chunk it has finished).  Highest chunk first.
        const int S_c = pipe_chunk_steps();
        int* flagsA = F.flags;
        int* flagsB = F.flags + static_cast<size_t>(S_c) * 64;
        cudaStream_t sr = pipe_prio() ? pp->sA : st;                 // layer 1's BPTT stream
        cudaEventRecord(pp->ev0, st);
        if (sr != st) cudaStreamWaitEvent(sr, pp->ev0, 0);
        cudaStreamWaitEvent(pp->sB, pp->ev0, 0);
        cudaStreamWaitEvent(pp->sC, pp->ev0, 0);
        const int n_chunks = (n.T + S_c - 1) / S_c;
        for (int c = n_chunks - 1; c >= 0; --c) {
            const int t0 = c * S_c, t1 = t0 + S_c < n.T ? t0 + S_c : n.T;
            const long long r0 = static_cast<long long>(t0) * n.B, rows = static_cast<long long>(t1 - t0) * n.B;
            FT_TRY(launch_lstm_bwd_chunk(n.T, n.B, t0, t1, F.dh, H, F.w.w_hh1, S_.gates1, S_.c1, out_lens, F.dG1, F.dcarry1, flagsA, sr));
            cudaEventRecord(pp->evA, sr);
            cudaStreamWaitEvent(pp->sC, pp->evA, 0);
            FT_TRY(gemm_dgrad(pp->sC, rows, H, G, F.dG1 + r0 * G, G, F.w.w_ih1, H, 0, F.dh + r0 * H, H, nullptr, 0, nullptr, 0));
            cudaEventRecord(pp->evG, pp->sC);
            cudaStreamWaitEvent(pp->sB, pp->evG, 0);
            FT_TRY(launch_lstm_bwd_chunk(n.T, n.B, t0, t1, F.dh, H, F.w.w_hh0, S_.gates0, S_.c0, out_lens, F.dG0, F.dcarry0, flagsB, pp->sB));
            if (attb_pipe && S_c % 64 == 0) {
                if (c == n_chunks - 1) cudaStreamWaitEvent(pp->sD, pp->ev0, 0);
                cudaEventRecord(pp->evD, pp->sB);
                cudaStreamWaitEvent(pp->sD, pp->evD, 0);
                FT_TRY(attention_bwd_rows(pp->sD, t0, t1));
                attention_bwd_done = true;
            }
        }
        if (attention_bwd_done) { cudaEventRecord(pp->evD, pp->sD); cudaStreamWaitEvent(st, pp->evD, 0); }
        // layer-1 weight gradients run under the tail of layer 0 (side stream forks from st = all of layer 1 done)
        if (sr != st) { cudaEventRecord(pp->evJ, sr); cudaStreamWaitEvent(st, pp->evJ, 0); }
        ss = fork_side(sd, st);
        FT_TRY(gemm_wgrad(ss, G, H, Rm, F.dG1 + static_cast<size_t>(n.B) * G, G, S_.h1_16, H, g.lstm_w_hh1, H, iS));
        FT_TRY(gemm_wgrad(ss, G, H, R, F.dG1, G, S_.h0_16, H, g.lstm_w_ih1, H, iS));
        FT_TRY(launch_colsum(F.dG1, 0, G, R, G, g.lstm_b_ih1, iS, ss));
        FT_TRY(copy_f32(g.lstm_b_hh1, g.lstm_b_ih1, G, ss));
        cudaEventRecord(pp->evB, pp->sB);
        cudaStreamWaitEvent(st, pp->evB, 0);
    } else {
    // 5. lstm layer 1
    FT_TRY(launch_lstm_bwd(n.T, n.B, F.dh, H, F.w.w_hh1, S_.gates1, S_.c1, out_lens, F.dG1, F.flags, st));
    ss = fork_side(sd, st);
    FT_TRY(gemm_wgrad(ss, G, H, Rm, F.dG1 + static_cast<size_t>(n.B) * G, G, S_.h1_16, H, g.lstm_w_hh1, H, iS));
    FT_TRY(gemm_wgrad(ss, G, H, R, F.dG1, G, S_.h0_16, H, g.lstm_w_ih1, H, iS));
    FT_TRY(launch_colsum(F.dG1, 0, G, R, G, g.lstm_b_ih1, iS, ss));
    FT_TRY(copy_f32(g.lstm_b_hh1, g.lstm_b_ih1, G, ss));
    FT_TRY(gemm_dgrad(st, R, H, G, F.dG1, G, F.w.w_ih1, H, 0, F.dh, H, nullptr, 0, nullptr, 0));

    // 6. lstm layer 0
    FT_TRY(launch_lstm_bwd(n.T, n.B, F.dh, H, F.w.w_hh0, S_.gates0, S_.c0, out_lens, F.dG0, F.flags, st));
    }
    // (the weight gradients of lstm layer 0 and of the attention projections are launched by ar_step_bwd_attn, on the side
    //  stream underneath the attention LSTM's BPTT: dG0, dQ16, dK16, dV16 stay in the scratch area until then)
    // Attention backward / attention-LSTM BPTT overlap (the mirror image of the forward overlap; FT_ATT_OVERLAP_BWD = chunk length,
    // default 256, 0 disables): only when the caller lets this call run the attention LSTM's BPTT as well (ft_ar_step_bwd: flows
    // whose d_text nobody is waiting for).  Highest chunk first: the second stream prepares chunk c's dhA while the BPTT kernel
    // (64 SMs) works on chunk c + 1.
    static int attb_chunk = -1;
    if (attb_chunk < 0) { const char* e = getenv("FT_ATT_OVERLAP_BWD"); attb_chunk = e ? atoi(e) : 256; if (attb_chunk % 64) attb_chunk = 0; }
    Pipe* pq = (may_fuse_bptt && !attention_bwd_done && attb_chunk > 0 && F.dcarry1 && n.B <= 32 && n.T >= attb_chunk + 64) ? get_pipe(st) : nullptr;
    if (pq) {
        FT_TRY(launch_cast(w.attn_lstm_w_ih, 2, F.w.w_ih_a, 0, static_cast<long long>(G) * n.M, st));
        FT_TRY(launch_transpose_cast_f16(w.attn_lstm_w_hh, F.w.w_hh_a, G, H, st));
        cudaStream_t sr = pipe_prio() ? pq->sA : st;
        cudaEventRecord(pq->ev0, st);
        if (sr != st) cudaStreamWaitEvent(sr, pq->ev0, 0);
        cudaStreamWaitEvent(pq->sC, pq->ev0, 0);
        const int CH = attb_chunk, n_chunks = (n.T + CH - 1) / CH;
        for (int c = n_chunks - 1; c >= 0; --c) {
            const int t0 = c * CH, t1 = t0 + CH < n.T ? t0 + CH : n.T;
            FT_TRY(attention_bwd_rows(pq->sC, t0, t1));
            cudaEventRecord(pq->evG, pq->sC);
            cudaStreamWaitEvent(sr, pq->evG, 0);
            // 10. attention_lstm  (dhA = C.dd[:, 0:H], pitch D)
            FT_TRY(launch_lstm_bwd_chunk(n.T, n.B, t0, t1, C.dd, n.D, F.w.w_hh_a, S_.gatesA, S_.cA, out_lens, F.dGa, F.dcarry1, F.flags, sr));
        }
        cudaStreamWaitEvent(st, pq->evG, 0);          // all of sC (the last record) ...
        if (sr != st) { cudaEventRecord(pq->evJ, sr); cudaStreamWaitEvent(st, pq->evJ, 0); }     // ... and the BPTT
        if (bptt_done) *bptt_done = true;
    } else if (!attention_bwd_done) {
        FT_TRY(attention_bwd_rows(st, 0, n.T));
    }

    // 9b. key / value projections -> d_text
    FT_TRY(launch_cast(F.dK, 2, F.dK16, 0, RL * n.A, st));
    FT_TRY(launch_cast(F.dV, 2, F.dV16, 0, RL * n.A, st));
    FT_TRY(gemm_dgrad(st, RL, n.E, n.A, F.dK16, n.A, F.w.wk, n.E, 0, d_text, n.E, nullptr, 0, nullptr, 0, iS));
    FT_TRY(gemm_dgrad(st, RL, n.E, n.A, F.dV16, n.A, F.w.wv, n.E, 1, d_text, n.E, nullptr, 0, nullptr, 0, iS));

    join_side(sd, st);                   // everything this call launched is ordered before what the caller enqueues next
    return 0;
}

// Attention-LSTM part: BPTT over dhA (carry), its weight gradients, and the input gradient d_mel (coupling path from carry +
// the shifted attention-LSTM path, back to natural time, loss scale undone).
int ar_step_bwd_attn(const FtArStepDesc& d, const FtArStepWeights& w, const int* out_lens, float* d_mel, const FtArStepWeights& g,
                     void* saved, void* scratch, void* carry, cudaStream_t st, bool bptt_done = false) {
    FT_TRY(check_desc(d));
    const Dims n(d);
    Plan ps; ps.base = static_cast<uint8_t*>(saved);
    Saved S_; S_.plan(ps, d);
    Plan pb; pb.base = static_cast<uint8_t*>(scratch);
    BwdScratch F; F.plan(pb, d);
    Plan pc; pc.base = static_cast<uint8_t*>(carry);
    BwdCarry C; C.plan(pc, d);
    const long long R = n.R, Rm = n.R - n.B;
    const float* iS = C.scale + 1;
    Side* sd = get_side(st);
    cudaStream_t ss;
    const long long RL = n.RL;
    if (!bptt_done) {                    // (a fused ar_step_bwd_main made these copies and ran the BPTT already)
        FT_TRY(launch_cast(w.attn_lstm_w_ih, 2, F.w.w_ih_a, 0, static_cast<long long>(G) * n.M, st));
        FT_TRY(launch_transpose_cast_f16(w.attn_lstm_w_hh, F.w.w_hh_a, G, H, st));
    }

    // deferred from ar_step_bwd_main: weight gradients of lstm layer 0 and of the attention projections, on the side stream,
    // underneath the BPTT kernel below (64 of the 148 SMs)
    ss = fork_side(sd, st);
    FT_TRY(gemm_wgrad(ss, G, H, Rm, F.dG0 + static_cast<size_t>(n.B) * G, G, S_.h0_16, H, g.lstm_w_hh0, H, iS));
    FT_TRY(gemm_wgrad(ss, G, n.D, R, F.dG0, G, S_.d16, n.D, g.lstm_w_ih0, n.D, iS));
    FT_TRY(launch_colsum(F.dG0, 0, G, R, G, g.lstm_b_ih0, iS, ss));
    FT_TRY(copy_f32(g.lstm_b_hh0, g.lstm_b_ih0, G, ss));
    FT_TRY(gemm_wgrad(ss, n.A, H, R, F.dQ16, n.A, S_.d16, n.D, g.att_query, H, iS));
    FT_TRY(gemm_wgrad(ss, n.A, n.E, RL, F.dK16, n.A, S_.text16, n.E, g.att_key, n.E, iS));
    FT_TRY(gemm_wgrad(ss, n.A, n.E, RL, F.dV16, n.A, S_.text16, n.E, g.att_value, n.E, iS));

    // 10. attention_lstm  (dhA = C.dd[:, 0:H], pitch D)
    if (!bptt_done) FT_TRY(launch_lstm_bwd(n.T, n.B, C.dd, n.D, F.w.w_hh_a, S_.gatesA, S_.cA, out_lens, F.dGa, F.flags, st));
    ss = fork_side(sd, st);
    FT_TRY(gemm_wgrad(ss, G, H, Rm, F.dGa + static_cast<size_t>(n.B) * G, G, S_.d16, n.D, g.attn_lstm_w_hh, H, iS));
    FT_TRY(gemm_wgrad(ss, G, n.M, R, F.dGa, G, S_.mel_in16, n.M, g.attn_lstm_w_ih, n.M, iS));
    FT_TRY(launch_colsum(F.dGa, 0, G, R, G, g.attn_lstm_b_ih, iS, ss));
    FT_TRY(copy_f32(g.attn_lstm_b_hh, g.attn_lstm_b_ih, G, ss));
    FT_TRY(gemm_dgrad(st, R, n.M, G, F.dGa, G, F.w.w_ih_a, n.M, 0, F.dmel_in, n.M, nullptr, 0, nullptr, 0));

    // 11. input gradient: coupling path + (shifted) attention_lstm path, back to natural time, loss scale undone
    if (d_mel) FT_TRY(launch_combine_dmel(C.dmel_flow, F.dmel_in, out_lens, n.T, n.B, n.M, d.reversed, d_mel, iS, st));
    join_side(sd, st);                   // everything this call launched is ordered before what the caller enqueues next
    return 0;
}

}  // namespace ft

// =================================================================================================== C ABI
extern "C" {

size_t ft_ar_step_saved_bytes(const FtArStepDesc* d) {
    ft::Plan p; ft::Saved s; s.plan(p, *d);
    return p.off + 256;
}
size_t ft_ar_step_bwd_carry_bytes(const FtArStepDesc* d) {
    ft::Plan pc; ft::BwdCarry c; c.plan(pc, *d);
    return pc.off + 256;
}
size_t ft_ar_step_scratch_bytes(const FtArStepDesc* d) {
    ft::Plan pf; ft::FwdScratch f; f.plan(pf, *d);
    ft::Plan pb; ft::BwdScratch b; b.plan(pb, *d);
    // the one-call backward (ft_ar_step_bwd) places its carry area behind the backward scratch
    const size_t bwd = ((pb.off + 255) & ~static_cast<size_t>(255)) + ft_ar_step_bwd_carry_bytes(d);
    return (pf.off > bwd ? pf.off : bwd) + 256;
}
int ft_ar_step_saved_lookup(const FtArStepDesc* d, const char* name, size_t* offset, size_t* bytes) {
    std::vector<ft::Region> regs;
    ft::Plan p; p.regs = &regs;
    ft::Saved s; s.plan(p, *d);
    for (const auto& r : regs)
        if (std::strcmp(r.name, name) == 0) { *offset = r.off; *bytes = r.bytes; return 0; }
    return ft::ft_set_error("ft_ar_step_saved_lookup: unknown region");
}

void ft_ar_step_set_text_ready_event(void* cuda_event) { ft::set_text_ready_event(cuda_event); }

int ft_ar_step_fwd(const FtArStepDesc* d, const FtArStepWeights* w, const float* mel, const float* text,
                   const int* in_lens, const int* out_lens, const float* attn_prior, float* mel_out, float* log_s,
                   float* gates, float* attn, float* attn_logprob, void* saved, void* scratch, void* stream) {
    if (!d || !w || !mel || !text || !mel_out || !log_s || !attn || !attn_logprob || !saved || !scratch)
        return ft::ft_set_error("ft_ar_step_fwd: NULL argument");
    return ft::ar_step_fwd(*d, *w, mel, text, in_lens, out_lens, attn_prior, mel_out, log_s, gates, attn, attn_logprob,
                           saved, scratch, static_cast<cudaStream_t>(stream));
}

int ft_ar_step_bwd(const FtArStepDesc* d, const FtArStepWeights* w, const float* mel, const int* in_lens,
                   const int* out_lens, const float* attn, const float* d_mel_out, const float* d_log_s,
                   const float* d_gates, const float* d_attn, const float* d_attn_logprob, float* d_mel, float* d_text,
                   const FtArStepWeights* g, void* saved, void* scratch, void* stream) {
    if (!d || !w || !mel || !attn || !d_text || !g || !saved || !scratch) return ft::ft_set_error("ft_ar_step_bwd: NULL argument");
    ft::Plan pb; ft::BwdScratch b; b.plan(pb, *d);
    void* carry = static_cast<uint8_t*>(scratch) + ((pb.off + 255) & ~static_cast<size_t>(255));
    bool bptt_done = false;           // the one-call form may overlap the attention backward with the attention LSTM's BPTT
    if (ft::ar_step_bwd_main(*d, *w, mel, in_lens, out_lens, attn, d_mel_out, d_log_s, d_gates, d_attn, d_attn_logprob, d_text, *g,
                             saved, scratch, carry, static_cast<cudaStream_t>(stream), true, &bptt_done) != 0) return -1;
    return ft::ar_step_bwd_attn(*d, *w, out_lens, d_mel, *g, saved, scratch, carry, static_cast<cudaStream_t>(stream), bptt_done);
}

int ft_ar_step_bwd_main(const FtArStepDesc* d, const FtArStepWeights* w, const float* mel, const int* in_lens,
                        const int* out_lens, const float* attn, const float* d_mel_out, const float* d_log_s,
                        const float* d_gates, const float* d_attn, const float* d_attn_logprob, float* d_text,
                        const FtArStepWeights* g, void* saved, void* scratch, void* carry, void* stream) {
    if (!d || !w || !mel || !attn || !d_text || !g || !saved || !scratch || !carry) return ft::ft_set_error("ft_ar_step_bwd_main: NULL argument");
    return ft::ar_step_bwd_main(*d, *w, mel, in_lens, out_lens, attn, d_mel_out, d_log_s, d_gates, d_attn, d_attn_logprob, d_text,
                                *g, saved, scratch, carry, static_cast<cudaStream_t>(stream));
}

int ft_ar_step_bwd_attn_lstm(const FtArStepDesc* d, const FtArStepWeights* w, const int* out_lens, float* d_mel,
                             const FtArStepWeights* g, void* saved, void* scratch, void* carry, void* stream) {
    if (!d || !w || !g || !saved || !scratch || !carry) return ft::ft_set_error("ft_ar_step_bwd_attn_lstm: NULL argument");
    return ft::ar_step_bwd_attn(*d, *w, out_lens, d_mel, *g, saved, scratch, carry, static_cast<cudaStream_t>(stream));
}

int ft_nll_reduce(const float* z, const float* const* log_s_list, int n_flows, const float* gate,
                  const float* gate_target, const int* out_lens, int T, int B, int M, float* sums, void* stream) {
    return ft::launch_nll_reduce(z, log_s_list, n_flows, gate, gate_target, out_lens, T, B, M, sums,
                                 static_cast<cudaStream_t>(stream));
}
int ft_nll_grad(const float* z, const float* gate, const float* gate_target, const int* out_lens, int T, int B, int M,
                float sigma, const float* sums, const float* g_nll, const float* g_gate, float* dz, float* dlog_s,
                float* dgate, void* stream) {
    return ft::launch_nll_grad(z, gate, gate_target, out_lens, T, B, M, sigma, sums, g_nll, g_gate, dz, dlog_s, dgate,
                               static_cast<cudaStream_t>(stream));
}

}  // extern "C"
