// C ABI of libdcgru_b200 (declared in include/dcgru_b200.h): argument checks, shared-memory /
// tiling plans, workspace carving and kernel launches.  No device memory is allocated here.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "dw.cuh"
#include "dcgru_b200.h"

using namespace dcgru;

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess) return fail("%s: %s", #expr, cudaGetErrorString(e_));     \
    } while (0)

// ---- optional per-kernel device timing (bench.py's roofline leg) -------------------------------------
// When enabled, every kernel launch below is bracketed by two cudaEvents recorded on the launching
// stream; dcgru_timing_collect() synchronises them and reports count / total ms per kernel name.
struct TimingRec { const char* name; cudaEvent_t a, b; };
std::mutex g_tmu;
bool g_timing = false;
std::vector<TimingRec> g_recs;
struct Timed {
    cudaStream_t st; cudaEvent_t b = nullptr;
    Timed(const char* name, cudaStream_t s) : st(s) {
        if (!g_timing) return;
        std::lock_guard<std::mutex> lk(g_tmu);
        TimingRec r; r.name = name;
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
        cudaEventRecord(r.a, st);
        b = r.b;
        g_recs.push_back(r);
    }
    ~Timed() { if (b) cudaEventRecord(b, st); }
};
#define LAUNCH(name, expr)          \
    do {                            \
        Timed t_(name, st);         \
        CUDA_TRY(expr);             \
    } while (0)

struct DevInfo { int sms = 0; int smem = 0; };
DevInfo query_dev(int dev) {
    DevInfo r;
    cudaDeviceGetAttribute(&r.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&r.smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return r;
}
// per-device attributes, filled once per device under std::call_once (callers: training thread + autograd thread)
const DevInfo& devinfo() {
    static DevInfo d[64];
    static std::once_flag once[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::call_once(once[dev], [dev] { d[dev] = query_dev(dev); });
    return d[dev];
}

// ---- side stream of the backward calls ---------------------------------------------------------------------------------
// The bias gradient (column sums of the dA image) is independent of the dX and dW GEMMs that follow the recurrent BPTT
// kernel, memory-bound and small: it is forked onto a library-owned non-blocking stream right after rnn_bwd and joined
// before the call returns, so it shares the SMs (its CTAs fit next to bulk_dp / dw_mm16) and the spare HBM bandwidth with
// them.  Plain event fork / join on the caller's stream: legal under CUDA-graph capture, no host synchronisation.  One
// stream + two events per (host thread, device), created on first use and kept for the life of the process (the only
// resources the library owns); DCGRU_SIDE_STREAM=0 and the per-kernel timing mode keep everything on the caller's stream.
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream* side_stream() {
    static thread_local SideStream tab[64];
    { const char* e = getenv("DCGRU_SIDE_STREAM"); if (e && e[0] == '0') return nullptr; }
    if (g_timing) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideStream& t = tab[dev];
    if (!t.s) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        (void)cs;
        if (cudaStreamCreateWithFlags(&t.s, cudaStreamNonBlocking) != cudaSuccess) { t.s = nullptr; return nullptr; }
        if (cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return (t.fork && t.join) ? &t : nullptr;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int check_desc(const dcgru_cell_desc* d) {
    if (!d) return fail("null descriptor");
    if (d->num_nodes < 1 || d->num_nodes > NP) return fail("num_nodes=%d unsupported (1..%d)", d->num_nodes, NP);
    if (d->hid_dim != 32 && d->hid_dim != 64 && d->hid_dim != 128)
        return fail("hid_dim=%d unsupported (32, 64 or 128)", d->hid_dim);
    if (d->input_dim < 4 || d->input_dim % 4) return fail("input_dim=%d must be a positive multiple of 4", d->input_dim);
    if (d->num_supports < 1 || d->num_supports > 2) return fail("num_supports=%d unsupported (1 or 2)", d->num_supports);
    if (d->max_diffusion_step < 0 || d->max_diffusion_step > 3)
        return fail("max_diffusion_step=%d unsupported (0..3)", d->max_diffusion_step);
    if (d->activation != DCGRU_ACT_TANH && d->activation != DCGRU_ACT_RELU) return fail("bad activation");
    return 0;
}
inline int Mof(const dcgru_cell_desc* d) { return d->num_supports * d->max_diffusion_step + 1; }

const double kCost[9] = {0, 1.6, 1.25, 0, 1.0, 0, 0, 0, 0.95};

// forward plan: samples per CTA, K-chunk, shared memory
struct FwdPlan { int SB = 0, KC = 0, smem = 0, FoPad = 0; };
bool plan_fwd(int H, int cmax, int M, int B, int Fo /*0 = encoder*/, FwdPlan* out) {
    const DevInfo& di = devinfo();
    double best = 1e30;
    for (int SB = 1; SB <= 8; SB *= 2) {
        int hs = H * SB;
        if (hs != 64 && hs != 128 && hs != 256) continue;
        int fopad = 0;
        if (Fo > 0) { int q = (64 / SB) * 4; fopad = (Fo + q - 1) / q * q; }
        for (int KC = 16; KC >= 2; KC /= 2) {
            FwdLayout L = fwd_layout(SB, H, cmax, M, KC);
            if (L.total * 4 > di.smem) continue;
            if (Fo > 0 && H * fopad > 2 * L.wbuf) continue;
            long ctas = (B + SB - 1) / SB;
            long waves = (ctas + di.sms - 1) / di.sms;
            double cost = waves * SB * kCost[SB];
            if (cost < best) { best = cost; out->SB = SB; out->KC = KC; out->smem = L.total * 4; out->FoPad = fopad; }
            break;   // largest KC that fits for this SB
        }
    }
    return best < 1e30;
}
struct BwdPlan { int SB = 0, smem = 0; };
bool plan_bwd(int H, int M, int B, int Fo, BwdPlan* out) {
    const DevInfo& di = devinfo();
    double best = 1e30;
    for (int SB = 1; SB <= 8; SB *= 2) {
        BwdLayout L = bwd_layout(SB, H, M, Fo);
        if (L.total * 4 > di.smem) continue;
        if (bwd_zcols(L.kb, M) < 4) continue;
        long ctas = (B + SB - 1) / SB;
        long waves = (ctas + di.sms - 1) / di.sms;
        double cost = waves * SB * kCost[SB];
        if (cost < best) { best = cost; out->SB = SB; out->smem = L.total * 4; }
    }
    return best < 1e30;
}

// ---- weight-gradient job tables --------------------------------------------------------------------
int otile(int n) { return n <= 192 ? n : (n % 192 == 0 ? 192 : 128); }   // 3H/2H/H for H in {32,64,128}

int build_cell_jobs(int fin, int H, int M, DwJob* jobs) {
    int n = 0;
    const int zc = 64 / M;
    auto add = [&](int type, int zlen, int kkbase, int o_begin, int o_len) {
        int ot = otile(o_len);
        for (int o0 = 0; o0 < o_len; o0 += ot)
            for (int z0 = 0; z0 < zlen; z0 += zc) {
                DwJob j;
                j.type = type; j.z0 = z0; j.nz = (zlen - z0 < zc) ? zlen - z0 : zc;
                j.kk0 = kkbase + z0 * M; j.o0 = o_begin + o0; j.nco = ot;
                if (n < DW_MAXJOBS) jobs[n] = j;
                ++n;
            }
    };
    add(0, fin, 0, 0, 3 * H);
    add(1, H, fin * M, 0, 2 * H);
    add(2, H, fin * M, 2 * H, H);
    return n;
}
int build_proj_jobs(int Fo, int H, DwJob* jobs) {
    int n = 0;
    int ot = otile(H);
    for (int o0 = 0; o0 < H; o0 += ot)
        for (int z0 = 0; z0 < Fo; z0 += 64) {
            DwJob j;
            j.type = 3; j.z0 = z0; j.nz = (Fo - z0 < 64) ? Fo - z0 : 64; j.kk0 = z0; j.o0 = o0; j.nco = ot;
            if (n < DW_MAXJOBS) jobs[n] = j;
            ++n;
        }
    return n;
}
int dw_nsplit(int B, int njobs) {
    const DevInfo& di = devinfo();
    int ngroups = (B + 3) / 4;
    int want = (4 * di.sms + njobs - 1) / njobs;
    if (want < 1) want = 1;
    return ngroups < want ? ngroups : want;
}

// ---- tensor-core weight-gradient plan (dw_tc.cu) ------------------------------------------------------
bool tc_enabled() {
    const char* e = getenv("DCGRU_DISABLE_TC");      // read per call so tests can compare both paths
    return !(e && e[0] == '1');
}
// jobs for dw_tc_kernel: 128-row slabs of W rows; DwJob.nz = rows in the slab, DwJob.z0 = carries the db row
int build_cell_jobs_tc(int fin, int H, int M, DwJob* jobs) {
    int n = 0;
    auto add = [&](int type, int kbegin, int kend, int o_begin, int o_len) {
        int ot = otile(o_len);
        for (int o0 = 0; o0 < o_len; o0 += ot) {
            bool have_db = (type != 0);
            for (int k0 = kbegin; k0 < kend; k0 += 128) {
                DwJob j;
                j.type = type; j.kk0 = k0; j.nz = (kend - k0 < 128) ? kend - k0 : 128; j.o0 = o_begin + o0; j.nco = ot;
                j.z0 = 0;
                if (!have_db && j.nz < 128) { j.z0 = 1; have_db = true; }
                if (n < DW_MAXJOBS) jobs[n] = j;
                ++n;
            }
            if (!have_db) {                  // no slab has a spare row: an empty slab that only carries db
                DwJob j;
                j.type = 0; j.kk0 = 0; j.nz = 0; j.o0 = o_begin + o0; j.nco = ot; j.z0 = 1;
                if (n < DW_MAXJOBS) jobs[n] = j;
                ++n;
            }
        }
    };
    add(0, 0, fin * M, 0, 3 * H);
    add(1, fin * M, (fin + H) * M, 0, 2 * H);
    add(2, fin * M, (fin + H) * M, 2 * H, H);
    return n;
}
struct CellDwPlan { bool tc; int njobs, nsplit, nco_max; };
CellDwPlan plan_cell_dw(int fin, int H, int M, int B, int T, DwJob* jobs) {
    DwJob tmp[DW_MAXJOBS];
    if (!jobs) jobs = tmp;
    CellDwPlan pl;
    const DevInfo& di = devinfo();
    pl.nco_max = otile(3 * H);
    pl.tc = tc_enabled() && M >= 3 && (H == 64 || H == 128) && di.sms > 0 &&
            dw_tc_smem_bytes(M, pl.nco_max) + 1088 <= di.smem;   // + static smem (mbarriers, 1 KB alignment)
    if (pl.tc) {
        pl.njobs = build_cell_jobs_tc(fin, H, M, jobs);
        if (pl.njobs > DW_MAXJOBS) pl.tc = false;
    }
    if (pl.tc) {
        long nchunk = ((long)T * B + 1) / 2;
        long want = di.sms / pl.njobs;           // one resident CTA per SM: exactly one wave
        if (want < 1) want = 1;
        pl.nsplit = (int)(nchunk < want ? nchunk : want);
    } else {
        pl.njobs = build_cell_jobs(fin, H, M, jobs);
        pl.nsplit = dw_nsplit(B, pl.njobs);
    }
    return pl;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
    char* base; size_t off = 0, cap;
    Carver(void* b, size_t c) : base((char*)b), cap(c) {}
    float* take(size_t nfloats) {
        size_t o = off;
        off = align_up(off + nfloats * 4);
        return (float*)(base + o);
    }
};

}  // namespace

extern "C" {

int dcgru_version(void) { return 100; }

int dcgru_tc_selftest(const float* A, const float* B, float* C, int32_t N, int32_t K, void* stream) {
    if (!A || !B || !C) return fail("null pointer");
    if (K < 32 || K % 32) return fail("K=%d must be a positive multiple of 32", K);
    if (N != 64 && N != 128 && N != 192 && N != 256) return fail("N=%d unsupported", N);
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("tc_selftest", launch_tc_selftest(A, B, C, N, K, st));
    return 0;
}

int dcgru_tc_probe(const float* a_img, int32_t a_bytes, const float* b_img, int32_t b_bytes, uint32_t a_lbo,
                   uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int32_t a_mn_major, int32_t b_mn_major,
                   int32_t a_layout_type, int32_t b_layout_type, float* D, int32_t N, void* stream) {
    if (!a_img || !b_img || !D) return fail("null pointer");
    if (a_bytes < 16 || a_bytes > 16384 || b_bytes < 16 || b_bytes > 16384 || a_bytes % 4 || b_bytes % 4)
        return fail("image sizes must be 16..16384 bytes");
    if (N < 16 || N > 256 || N % 16) return fail("N=%d unsupported", N);
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24) |
                           (a_mn_major ? 1u << 15 : 0u) | (b_mn_major ? 1u << 16 : 0u);
    LAUNCH("tc_probe", launch_tc_probe(a_img, a_bytes, b_img, b_bytes, a_lbo, a_sbo, b_lbo, b_sbo, idesc,
                                       (uint32_t)a_layout_type & 7u, (uint32_t)b_layout_type & 7u, D, N, st));
    return 0;
}

int dcgru_timing_enable(int on) {
    std::lock_guard<std::mutex> lk(g_tmu);
    for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_recs.clear();
    g_timing = on != 0;
    return 0;
}

int dcgru_timing_collect(char* buf, size_t cap) {
    std::lock_guard<std::mutex> lk(g_tmu);
    std::map<std::string, std::pair<long, double>> agg;
    for (auto& r : g_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess)
            return fail("timing events not complete");
        auto& e = agg[r.name];
        e.first += 1; e.second += ms;
    }
    std::string out;
    char line[160];
    for (auto& kv : agg) {
        snprintf(line, sizeof line, "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (!buf || out.size() + 1 > cap) return fail("timing buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}
const char* dcgru_last_error(void) { return g_err.c_str(); }

int dcgru_graph_poly(int32_t batch, int32_t num_nodes, int32_t max_diffusion_step, int32_t num_supports,
                     const float* const* supports, const int64_t* support_bstride, float* P, void* stream) {
    if (batch < 1) return fail("batch < 1");
    if (num_nodes < 1 || num_nodes > NP) return fail("num_nodes=%d unsupported", num_nodes);
    if (num_supports < 1 || num_supports > 2) return fail("num_supports=%d unsupported", num_supports);
    if (max_diffusion_step < 0 || max_diffusion_step > 3) return fail("max_diffusion_step unsupported");
    long long bs[2] = {0, 0};
    for (int s = 0; s < num_supports; ++s) {
        if (!supports || !supports[s]) return fail("null support %d", s);
        bs[s] = support_bstride ? support_bstride[s] : (long long)num_nodes * num_nodes;
    }
    if (max_diffusion_step > 0 && !P) return fail("null P");
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("graph_poly", launch_graph_poly(batch, num_nodes, max_diffusion_step, num_supports, supports, bs, P, st));
    return 0;
}

int dcgru_corr_supports(int32_t batch, int32_t seq_len, int32_t num_nodes, int32_t feat, const float* clip,
                        int64_t stride_b, int64_t stride_t, float scale, float shift, int32_t top_k,
                        float* adj, float* support0, float* support1, void* stream) {
    if (batch < 1 || seq_len < 1 || feat < 1) return fail("empty clip");
    if (num_nodes < 2 || num_nodes > NP) return fail("num_nodes=%d unsupported", num_nodes);
    if (top_k < 0) return fail("top_k < 0 (the reference raises ValueError for top_k=None)");
    if (!clip || !support0 || !support1) return fail("null pointer");
    if ((size_t)num_nodes * (feat | 1) * 8 > (size_t)devinfo().smem) return fail("feature dim too large");
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("corr_supports", launch_corr_supports(batch, seq_len, num_nodes, feat, clip, stride_b, stride_t, scale,
                                                 shift, top_k, adj, support0, support1, st));
    return 0;
}

int dcgru_fft_features(int32_t batch, int32_t num_nodes, int32_t seq_len, int32_t window, const float* signal,
                       int64_t stride_b, int64_t stride_n, const int32_t* dest_channel, const float* log_scale,
                       const float* mean, const float* std, int32_t stat_len, float* raw, float* x, void* stream) {
    if (batch < 1 || seq_len < 1) return fail("empty clip");
    if (num_nodes < 1 || num_nodes > 32) return fail("num_nodes=%d unsupported (1..32)", num_nodes);
    if (window != 200) return fail("window=%d unsupported (FREQUENCY * time_step_size = 200 samples)", window);
    if (!signal) return fail("null signal");
    if (!raw && !x) return fail("no output requested");
    if (stat_len != 0 && stat_len != 1 && stat_len != num_nodes) return fail("stat_len=%d must be 0, 1 or num_nodes", stat_len);
    if (stat_len && (!mean || !std)) return fail("null mean/std");
    if ((raw && !aligned16(raw)) || (x && !aligned16(x))) return fail("outputs must be 16-byte aligned");
    if (!aligned16(signal) || stride_b % 4 || stride_n % 4)
        return fail("signal must be 16-byte aligned with batch / channel strides that are multiples of 4 samples (bulk copies)");
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("fft_features", launch_fft_features(batch, num_nodes, seq_len, signal, stride_b, stride_n, dest_channel, log_scale,
                                               mean, std, stat_len, raw, x, devinfo().sms, st));
    return 0;
}

// second-generation path (2xFP16: bulk_dp.cu, rnn_fwd.cu, rnn_bwd.cu, dw_mm16.cu) is the default for H = 64 cells;
// DCGRU_G2=0 selects the first-generation 3xTF32 kernels (H = 64, M = 3 only), DCGRU_DISABLE_TC=1 the fp32 FMA kernels
static bool g2_enabled() {
    const char* e = getenv("DCGRU_G2");
    return tc_enabled() && !(e && e[0] == '0');
}
static bool g2_fwd_supported(const dcgru_cell_desc* d) {
    const int M = Mof(d);
    const DevInfo& di = devinfo();
    return g2_enabled() && rnn_fwd_supported(d->num_nodes, d->hid_dim, M, di.smem) &&
           bulk_dp_supported(d->num_nodes, d->input_dim, M, 3 * d->hid_dim, false, di.smem);
}
struct G2FwdWs { size_t off_wx, off_wh, off_bias, off_xp, total; };
static G2FwdWs g2_fwd_ws(const dcgru_cell_desc* d, int B, int T) {
    G2FwdWs w;
    const int M = Mof(d);
    size_t o = 0;
    w.off_wx = o; o = align_up(o + bulk_wimg_bytes(d->input_dim, M, 3 * d->hid_dim));
    w.off_wh = o; o = align_up(o + rnn_fwd_wimg_bytes(M));
    w.off_bias = o; o = align_up(o + (size_t)3 * d->hid_dim * 4);
    w.off_xp = o; o = align_up(o + (size_t)T * B * d->num_nodes * 3 * d->hid_dim * 4);
    w.total = o;
    return w;
}

static bool g2_bwd_supported(const dcgru_cell_desc* d) {
    return g2_fwd_supported(d) && rnn_bwd_supported(d->num_nodes, d->hid_dim, Mof(d), devinfo().smem);
}
struct G2BwdWs { size_t off_wb, off_wdx, off_img, off_scale, off_part, off_cs, total; };
static G2BwdWs g2_bwd_ws(const dcgru_cell_desc* d, int B, int T) {
    G2BwdWs w;
    const int M = Mof(d), H = d->hid_dim;
    size_t o = 0;
    w.off_wb = o; o = align_up(o + rnn_bwd_wimg_bytes(M));
    w.off_wdx = o; o = align_up(o + bulk_wimg_bytes(3 * H, M, d->input_dim));
    w.off_img = o; o = align_up(o + g16_image_bytes(B, T, 3 * H));
    w.off_scale = o; o = align_up(o + 256);
    w.off_part = o; o = align_up(o + dw_mm16_part_floats(devinfo().sms) * 4);
    w.off_cs = o; o = align_up(o + colsum16_part_floats(H) * 4);
    w.total = o;
    return w;
}
static bool g2_gsave_enabled() {
    const char* e = getenv("DCGRU_DISABLE_GSAVE");
    return !(e && e[0] == '1');
}

// operand image of the tensor-core forward kernel (0: the configuration has no such path)
static size_t gsave_bytes_for(const dcgru_cell_desc* d, int B, int T) {
    if (!tc_enabled()) return 0;
    if (g2_fwd_supported(d)) {                  // second generation: fp16 image [tile*T+t][hi|lo][96][KKP]
        if (!g2_bwd_supported(d) || !g2_gsave_enabled() || dw_mm16_smem_bytes() + 2048 > devinfo().smem) return 0;
        return g16_image_bytes(B, T, g16_kkp(d->input_dim, d->hid_dim, Mof(d)));
    }
    { const char* e = getenv("DCGRU_DISABLE_GSAVE"); if (e && e[0] == '1') return 0; }
    const int M = Mof(d);
    const DevInfo& di = devinfo();
    if (!seq_fwd_tc_supported(d->num_nodes, d->input_dim, d->hid_dim, M, di.smem)) return 0;
    if (!seq_bwd_tc_supported(d->num_nodes, d->hid_dim, M, di.smem)) return 0;
    if (dw_mm_smem_bytes() + 2048 > di.smem) return 0;
    DwmmParams q;
    if (!dwmm_plan(d->input_dim, d->hid_dim, M, seq_tc_nslab(B, T), di.sms, &q)) return 0;
    return seq_fwd_tc_gsave_bytes(B, T, d->input_dim);
}

size_t dcgru_encoder_layer_gsave_bytes(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len) {
    if (check_desc(d) || batch < 1 || seq_len < 1) return 0;
    return gsave_bytes_for(d, batch, seq_len);
}

size_t dcgru_encoder_layer_fwd_workspace(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len) {
    if (check_desc(d) || batch < 1 || seq_len < 1) return 0;
    size_t n = align_up(seq_fwd_tc_wimg_bytes(d->input_dim)) + 256 + 16384;   // + debug stamps (DCGRU_DBG & 4)
    if (g2_fwd_supported(d)) { const size_t g = g2_fwd_ws(d, batch, seq_len).total; if (g > n) n = g; }
    return n;
}

int dcgru_encoder_layer_fwd(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len, const float* x,
                            int64_t x_stride_t, int64_t x_stride_b, const float* h0, const float* P,
                            const dcgru_cell_params* w, float* h_seq, float* ruc, void* gsave,
                            size_t gsave_bytes, void* workspace, size_t workspace_bytes, void* stream) {
    if (check_desc(d)) return 1;
    if (batch < 1 || seq_len < 1) return fail("empty batch/sequence");
    if (!x || !h0 || !w || !h_seq || !w->Wg || !w->bg || !w->Wc || !w->bc) return fail("null pointer");
    const int M = Mof(d);
    if (M > 1 && !P) return fail("null P");
    if (!aligned16(x) || !aligned16(h0) || !aligned16(h_seq) || !aligned16(w->Wg) || !aligned16(w->Wc) ||
        (x_stride_t % 4) || (x_stride_b % 4))
        return fail("x/h0/h_seq/weights must be 16-byte aligned with strides multiple of 4 floats");
    cudaStream_t st = (cudaStream_t)stream;
    // second-generation tensor-core path (tcgen05, 2xFP16): hoisted x-part GEMM over all steps, then the recurrence
    if (g2_fwd_supported(d) && workspace && aligned16(workspace) &&
        workspace_bytes >= g2_fwd_ws(d, batch, seq_len).total) {
        const int kkp = g16_kkp(d->input_dim, d->hid_dim, M), kxp = g16_kxp(d->input_dim, M);
        if (gsave) {
            const size_t need = gsave_bytes_for(d, batch, seq_len);
            if (need == 0) return fail("this configuration has no operand image: pass gsave = NULL");
            if (gsave_bytes < need) return fail("gsave too small (%zu < %zu bytes)", gsave_bytes, need);
            if (!aligned16(gsave)) return fail("gsave must be 16-byte aligned");
        }
        const G2FwdWs ws = g2_fwd_ws(d, batch, seq_len);
        const DevInfo& di = devinfo();
        const int H = d->hid_dim, N = d->num_nodes, fin = d->input_dim;
        uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
        float* xp = reinterpret_cast<float*>(wsb + ws.off_xp);
        float* bias = reinterpret_cast<float*>(wsb + ws.off_bias);                       // [bg | bc]
        CUDA_TRY(cudaMemcpyAsync(bias, w->bg, (size_t)2 * H * 4, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(bias + 2 * H, w->bc, (size_t)H * 4, cudaMemcpyDeviceToDevice, st));
        LAUNCH("pack_w16", launch_pack_w16(w->Wg, w->Wc, fin, H, M, 0, 3 * H, g16_nq(fin, M), wsb + ws.off_wx, st));
        LAUNCH("xproj", launch_bulk_dp(batch, seq_len, N, fin, M, 3 * H, 0, x, x_stride_t, x_stride_b, nullptr, P, wsb + ws.off_wx, bias,
                                       xp, (long long)batch * N * 3 * H, (long long)N * 3 * H, 3 * H, 1.f, nullptr, gsave, kkp,
                                       0, di.sms, di.smem, st));
        LAUNCH("rnn_fwd", launch_rnn_fwd(batch, seq_len, N, fin, M, d->activation, xp, h0, P, w->Wg, w->Wc, wsb + ws.off_wh,
                                         h_seq, ruc, gsave, kkp, kxp, st));
        return 0;
    }
    // tensor-core path (tcgen05, 3xTF32): K=2 / one support / 64 units -- the reference's default cell
    if (tc_enabled() && workspace && aligned16(workspace) &&
        workspace_bytes >= align_up(seq_fwd_tc_wimg_bytes(d->input_dim)) + 16384 &&
        seq_fwd_tc_supported(d->num_nodes, d->input_dim, d->hid_dim, M, devinfo().smem)) {
        if (gsave) {
            const size_t need = gsave_bytes_for(d, batch, seq_len);
            if (need == 0) return fail("this configuration has no operand image: pass gsave = NULL");
            if (gsave_bytes < need) return fail("gsave too small (%zu < %zu bytes)", gsave_bytes, need);
            if (!aligned16(gsave)) return fail("gsave must be 16-byte aligned");
        }
        LAUNCH("seq_fwd_tc", launch_seq_fwd_tc(batch, seq_len, d->num_nodes, d->input_dim, d->activation, x,
                                               x_stride_t, x_stride_b, h0, P, w->Wg, w->bg, w->Wc, w->bc,
                                               (float*)workspace, h_seq, ruc, gsave, st));
        return 0;
    }
    if (gsave) return fail("gsave given but the tensor-core forward path is not available for this call");
    FwdPlan pl;
    if (!plan_fwd(d->hid_dim, d->input_dim + d->hid_dim, M, batch, 0, &pl))
        return fail("no tiling fits shared memory (input_dim=%d hid=%d M=%d)", d->input_dim, d->hid_dim, M);
    FwdParams p;
    memset(&p, 0, sizeof p);
    p.B = batch; p.T = seq_len; p.N = d->num_nodes; p.H = d->hid_dim; p.M = M; p.act = d->activation;
    p.ncell = 1; p.KC = pl.KC; p.mode = 0;
    p.cell[0] = CellW{w->Wg, w->bg, w->Wc, w->bc, d->input_dim};
    p.P = P; p.x = x; p.xs_t = x_stride_t; p.xs_b = x_stride_b; p.h0 = h0; p.hseq = h_seq; p.ruc = ruc;
    LAUNCH("seq_fwd", launch_seq_fwd(p, pl.SB, pl.smem, st));
    return 0;
}

static size_t enc_bwd_ws(const dcgru_cell_desc* d, int B, int T, bool carve, void* ws, float** WgT, float** WcT,
                         float** dA, float** part, float** partb, int* nsplit, int* njobs, DwJob* jobs,
                         float** ptbuf = nullptr, float** daimg = nullptr, float** mmpart = nullptr,
                         float** cspart = nullptr, float** g2base = nullptr) {
    const int H = d->hid_dim, M = Mof(d), CM = (d->input_dim + H) * M;
    CellDwPlan dp = plan_cell_dw(d->input_dim, H, M, B, T, jobs);
    int nj = dp.njobs, ns = dp.nsplit;
    Carver c(ws, 0);
    float* a = c.take((size_t)2 * H * CM);
    float* b = c.take((size_t)H * CM);
    float* e = c.take((size_t)T * B * d->num_nodes * 3 * H);
    float* f = c.take((size_t)ns * CM * 3 * H);
    float* g = c.take((size_t)ns * 3 * H);
    float* pt = c.take(((dw_tc_pt_floats(B, M) + 63) / 64) * 64 + 2 * (seq_bwd_tc_wimg_bytes() / 4 + 64));   // P^T + BPTT weight image + dX weight image
    // operand-image path (dw_mm.cu): dA image, per-CTA partials, column-sum partials
    float *im = nullptr, *mp = nullptr, *cp = nullptr;
    if (gsave_bytes_for(d, B, T) > 0 && !g2_bwd_supported(d)) {
        im = c.take(seq_bwd_tc_daimg_bytes(B, T) / 4);
        mp = c.take(dwmm_part_floats(devinfo().sms));
        cp = c.take(colsum_part_floats(H));
    }
    if (carve) {
        *WgT = a; *WcT = b; *dA = e; *part = f; *partb = g; *nsplit = ns; *njobs = nj; *ptbuf = pt;
        *daimg = im; *mmpart = mp; *cspart = cp;
    }
    // second-generation backward (rnn_bwd.cu): weight images, dA operand image, gradient scale
    if (g2_bwd_supported(d)) {
        float* g2 = c.take(g2_bwd_ws(d, B, T).total / 4);
        if (carve && g2base) *g2base = g2;
    }
    return c.off;
}

size_t dcgru_encoder_layer_bwd_workspace(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len) {
    if (check_desc(d) || batch < 1 || seq_len < 1) return 0;
    return enc_bwd_ws(d, batch, seq_len, false, nullptr, 0, 0, 0, 0, 0, 0, 0, 0);
}

int dcgru_debug_dwmm_stamps(long long* out, int32_t n) {
    CUDA_TRY(dwmm_read_dbg(out, n));
    return 0;
}
int dcgru_debug_rnn_fwd_stamps(long long* out, int32_t n) {
    CUDA_TRY(rnn_fwd_read_dbg(out, n));
    return 0;
}
int dcgru_debug_rnn_bwd_stamps(long long* out, int32_t n) {
    CUDA_TRY(rnn_bwd_read_dbg(out, n));
    return 0;
}

// plan of the weight-gradient GEMM as plain integers (host logic only: testable without a device):
// out[0] = number of sets, out[1] = K blocks, then per set 4 + 6*ntile ints:
//   ntile, TMEM columns, CTAs, TMA boxes, and per tile kg0, nkg, og0, ncol, tcol, soff
int dcgru_debug_dwmm_plan(int32_t input_dim, int32_t batch, int32_t seq_len, int32_t num_sms, int32_t* out, int32_t cap) {
    if (!out || cap < 2) return fail("bad arguments");
    DwmmParams q;
    if (!dwmm_plan(input_dim, 64, 3, seq_tc_nslab(batch, seq_len), num_sms, &q)) return fail("no plan for this shape");
    int n = 0;
    auto put = [&](long v) { if (n < cap) out[n] = (int32_t)v; ++n; };
    put(q.nset); put(q.nkb);
    for (int s = 0; s < q.nset; ++s) {
        const DwmmSet& S = q.set[s];
        put(S.ntile); put(S.ncoltot); put(S.ncta); put(S.nbox);
        for (int j = 0; j < S.ntile; ++j) {
            const DwmmTile& t = S.tile[j];
            put(t.kg0); put(t.nkg); put(t.og0); put(t.ncol); put(t.tcol); put(t.soff);
        }
    }
    if (n > cap) return fail("output buffer too small (%d > %d)", n, cap);
    return 0;
}

int dcgru_debug_encoder_bwd_offsets(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len, size_t* out3) {
    if (check_desc(d) || batch < 1 || seq_len < 1 || !out3) return fail("bad arguments");
    float *WgT, *WcT, *dA, *part, *partb, *ptbuf, *daimg, *mmpart, *cspart;
    int nsplit, njobs;
    DwJob jobs[DW_MAXJOBS];
    enc_bwd_ws(d, batch, seq_len, true, nullptr, &WgT, &WcT, &dA, &part, &partb, &nsplit, &njobs, jobs, &ptbuf,
               &daimg, &mmpart, &cspart);
    out3[0] = (size_t)((char*)dA - (char*)nullptr);
    out3[1] = daimg ? (size_t)((char*)daimg - (char*)nullptr) : 0;
    out3[2] = mmpart ? (size_t)((char*)mmpart - (char*)nullptr) : 0;
    return 0;
}

// d_hsel / sel_t: the sparse form of d_hseq (sample b's (N*H) slab belongs to step sel_t[b]; dcgru_cls_head_bwd writes it)
static size_t sel_dense_bytes(const dcgru_cell_desc* d, int B, int T) {
    return align_up((size_t)T * B * d->num_nodes * d->hid_dim * 4);
}
static int encoder_layer_bwd_impl(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len, const float* x,
                                  int64_t x_stride_t, int64_t x_stride_b, const float* h0, const float* P,
                                  const dcgru_cell_params* w, const float* h_seq, const float* ruc,
                                  const float* d_hseq, const float* d_hlast, const float* d_hsel, const int32_t* sel_t,
                                  float* dx, float* dh0,
                                  const dcgru_cell_grads* g, const void* gsave, size_t gsave_bytes,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    if (check_desc(d)) return 1;
    if (batch < 1 || seq_len < 1) return fail("empty batch/sequence");
    if (!x || !h0 || !w || !h_seq || !ruc || !dh0 || !g || !workspace) return fail("null pointer");
    if (!g->dWg || !g->dbg || !g->dWc || !g->dbc) return fail("null gradient pointer");
    const int M = Mof(d), H = d->hid_dim, fin = d->input_dim, CM = (fin + H) * M;
    if (M > 1 && !P) return fail("null P");
    if (!aligned16(workspace) || !aligned16(h_seq) || !aligned16(ruc)) return fail("unaligned pointer");
    const size_t base_ws = dcgru_encoder_layer_bwd_workspace(d, batch, seq_len);
    if ((d_hsel ? align_up(base_ws) + sel_dense_bytes(d, batch, seq_len) : base_ws) > workspace_bytes) return fail("workspace too small");
    if (d_hsel && d_hseq) return fail("pass either the dense d_hseq or the sparse d_hsel, not both");
    if (d_hsel && !aligned16(d_hsel)) return fail("unaligned d_hsel");
    cudaStream_t st = (cudaStream_t)stream;
    const bool g2_path = g2_bwd_supported(d) && (!dx || bulk_dp_supported(d->num_nodes, 3 * H, M, fin, true, devinfo().smem));
    if (d_hsel && !g2_path) {
        // the first-generation / fp32 kernels take a dense gradient: expand the slab behind the regular workspace
        float* dense = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + align_up(base_ws));
        LAUNCH("scatter_sel", launch_scatter_sel(batch, seq_len, d->num_nodes * H, d_hsel, sel_t, dense, st));
        d_hseq = dense; d_hsel = nullptr; sel_t = nullptr;
    }
    float *WgT, *WcT, *dA, *part, *partb, *ptbuf, *daimg, *mmpart, *cspart;
    int nsplit, njobs;
    DwParams q;
    memset(&q, 0, sizeof q);
    float* g2base = nullptr;
    enc_bwd_ws(d, batch, seq_len, true, workspace, &WgT, &WcT, &dA, &part, &partb, &nsplit, &njobs, q.jobs, &ptbuf,
               &daimg, &mmpart, &cspart, &g2base);
    DwmmParams mm;
    bool use_mm = false;
    if (gsave && !g2_bwd_supported(d)) {
        const size_t need = gsave_bytes_for(d, batch, seq_len);
        if (need == 0 || !daimg) return fail("this configuration has no operand image: pass gsave = NULL");
        if (gsave_bytes < need) return fail("gsave too small (%zu < %zu bytes)", gsave_bytes, need);
        if (!dwmm_plan(fin, H, M, seq_tc_nslab(batch, seq_len), devinfo().sms, &mm)) return fail("dw_mm plan failed");
        use_mm = true;
    }
    if (njobs > DW_MAXJOBS) return fail("too many weight-gradient jobs (%d)", njobs);
    if (!(g2_path && g2base)) {                                 // (only the fp32 / first-generation BPTT kernels read transposed weights)
        LAUNCH("transpose", launch_transpose(w->Wg, CM, 2 * H, WgT, CM, st));
        LAUNCH("transpose", launch_transpose(w->Wc, CM, H, WcT, CM, st));
    }
    BwdPlan pl;
    if (!plan_bwd(H, M, batch, 0, &pl)) return fail("no backward tiling fits shared memory");
    BwdParams p;
    memset(&p, 0, sizeof p);
    p.B = batch; p.T = seq_len; p.N = d->num_nodes; p.H = H; p.M = M; p.act = d->activation; p.ncell = 1; p.mode = 0;
    p.cell[0] = CellWT{WgT, WcT, fin};
    p.P = P; p.h0 = h0; p.hseq = h_seq; p.ruc = ruc; p.d_hseq = d_hseq; p.d_hlast = d_hlast;
    p.dx = dx; p.dh0 = dh0; p.dA = dA;
    if (g2_path && g2base) {
        if (gsave) {
            const size_t need = gsave_bytes_for(d, batch, seq_len);
            if (need == 0) return fail("this configuration has no operand image: pass gsave = NULL");
            if (gsave_bytes < need) return fail("gsave too small (%zu < %zu bytes)", gsave_bytes, need);
        }
        // second-generation BPTT (2xFP16): gradient scale -> recurrent kernel (dA operand image) -> bulk dX GEMM
        const G2BwdWs ws = g2_bwd_ws(d, batch, seq_len);
        const DevInfo& di = devinfo();
        const int N = d->num_nodes;
        uint8_t* wsb = reinterpret_cast<uint8_t*>(g2base);
        float* scale = reinterpret_cast<float*>(wsb + ws.off_scale);
        const size_t nh = (size_t)batch * N * H;
        LAUNCH("grad_scale", launch_grad_scale(d_hseq, d_hseq ? (size_t)seq_len * nh : 0, d_hlast, d_hlast ? nh : 0,
                                               d_hsel, d_hsel ? nh : 0, reinterpret_cast<unsigned*>(scale + 8), scale, st));
        LAUNCH("rnn_bwd", launch_rnn_bwd(batch, seq_len, N, fin, M, d->activation, h0, h_seq, ruc, P, w->Wg, w->Wc, d_hseq,
                                         d_hlast, d_hsel, sel_t, wsb + ws.off_wb, scale, dh0, wsb + ws.off_img, st));
        // db = column sums of the dA image: fused into the weight-gradient GEMM, whose CTAs stream those rows through shared
        // memory anyway (DCGRU_FUSE_DB=0: separate colsum16 kernel, on the side stream when there is one)
        bool fuse_db = gsave != nullptr;
        { const char* e = getenv("DCGRU_FUSE_DB"); if (e && e[0] == '0') fuse_db = false; }
        SideStream* side = (gsave && !fuse_db) ? side_stream() : nullptr;
        if (side) {
            // db on the side stream, concurrent with the dX / dW GEMMs below (joined before the call returns)
            CUDA_TRY(cudaEventRecord(side->fork, st));
            CUDA_TRY(cudaStreamWaitEvent(side->s, side->fork, 0));
            {
                cudaStream_t st = side->s;                                  // (LAUNCH times and launches on `st`)
                LAUNCH("colsum16", launch_colsum16(wsb + ws.off_img, batch, seq_len, H, reinterpret_cast<float*>(wsb + ws.off_cs), scale,
                                                   g->dbg, g->dbc, st));
            }
            CUDA_TRY(cudaEventRecord(side->join, side->s));
        }
        if (dx) {
            LAUNCH("pack_w16", launch_pack_w16(w->Wg, w->Wc, fin, H, M, 1, fin, g16_nq(3 * H, M), wsb + ws.off_wdx, st));
            LAUNCH("dx16", launch_bulk_dp(batch, seq_len, N, 3 * H, M, fin, 1, nullptr, 0, 0, wsb + ws.off_img, P, wsb + ws.off_wdx,
                                          nullptr, dx, (long long)batch * N * fin, (long long)N * fin, fin, 1.f, scale, nullptr, 0,
                                          0, di.sms, di.smem, st));
        }
        if (gsave) {
            // weight gradient: GEMM over the two fp16 operand images; bias gradient: column sums of the dA image
            LAUNCH("dw_mm16", launch_dw_mm16(fin, H, M, batch, seq_len, gsave, wsb + ws.off_img, reinterpret_cast<float*>(wsb + ws.off_part),
                                             scale, di.sms, g->dWg, g->dWc, st, fuse_db ? reinterpret_cast<float*>(wsb + ws.off_cs) : nullptr,
                                             g->dbg, g->dbc));
            if (fuse_db) return 0;
            if (side) CUDA_TRY(cudaStreamWaitEvent(st, side->join, 0));
            else
                LAUNCH("colsum16", launch_colsum16(wsb + ws.off_img, batch, seq_len, H, reinterpret_cast<float*>(wsb + ws.off_cs), scale,
                                                   g->dbg, g->dbc, st));
            return 0;
        }
        // no operand image: bridge to the first-generation (recompute) weight-gradient kernels through a row-major fp32 dA
        LAUNCH("img_to_rows", launch_img_to_rows(wsb + ws.off_img, batch, seq_len, N, 3 * H, scale, dA, 1, 1, 0, st));
    } else if (tc_enabled() && seq_bwd_tc_supported(d->num_nodes, H, M, devinfo().smem)) {
        // recurrent part on the tensor cores; the input gradient is not recurrent -> bulk pass over all steps
        float* wimg_b = ptbuf + ((dw_tc_pt_floats(batch, M) + 63) / 64) * 64;
        LAUNCH("seq_bwd_tc", launch_seq_bwd_tc(batch, seq_len, d->num_nodes, fin, d->activation, h0, h_seq, ruc, P,
                                               w->Wg, w->Wc, d_hseq, d_hlast, wimg_b, dh0, dA,
                                               use_mm ? (void*)daimg : nullptr, st));
        if (dx) {
            if (fin == 64) {
                float* wimg_x = wimg_b + seq_bwd_tc_wimg_bytes() / 4 + 64;
                LAUNCH("dx_tc", launch_dx_tc(batch, seq_len, d->num_nodes, P, w->Wg, w->Wc, dA, wimg_x, dx, st));
            } else {
                p.mode = 2;
                LAUNCH("dx", launch_seq_bwd(p, pl.SB, pl.smem, st));
            }
        }
    } else {
        LAUNCH("seq_bwd", launch_seq_bwd(p, pl.SB, pl.smem, st));
    }
    // bulk weight gradients
    if (use_mm) {
        // GEMM over the operand images left by the forward (G) and backward (dA) sequence kernels
        mm.G = reinterpret_cast<const uint8_t*>(gsave);
        mm.DA = reinterpret_cast<const uint8_t*>(daimg);
        mm.part = mmpart;
        LAUNCH("dw_mm", launch_dw_mm(mm, fin, H, M, g->dWg, g->dWc, st));
        LAUNCH("colsum", launch_colsum(dA, (long)seq_len * batch * d->num_nodes, H, cspart, g->dbg, g->dbc, st));
        return 0;
    }
    q.B = batch; q.T = seq_len; q.N = d->num_nodes; q.H = H; q.M = M; q.nsplit = nsplit; q.mode = 0;
    q.layer = 0; q.ncell = 1; q.fin = fin; q.Fo = 0;
    q.P = P; q.x = x; q.xs_t = x_stride_t; q.xs_b = x_stride_b; q.h0 = h0; q.hseq = h_seq; q.ruc = ruc; q.dA = dA;
    q.part = part; q.partb = partb;
    if (plan_cell_dw(fin, H, M, batch, seq_len, nullptr).tc) {
        LAUNCH("make_pt", launch_make_pt(P, batch, M, d->num_nodes, ptbuf, st));
        q.dY = ptbuf;                             // dw_tc_kernel reads the padded P^T through this field
        LAUNCH("dw_tc", launch_dw_tc(q, njobs, otile(3 * H), st));
    }
    else
        LAUNCH("dw", launch_dw(q, njobs, otile(3 * H), st));
    LAUNCH("reduce", launch_reduce_cell(part, partb, nsplit, CM, H, g->dWg, g->dbg, g->dWc, g->dbc, st));
    return 0;
}

int dcgru_encoder_layer_bwd(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len, const float* x,
                            int64_t x_stride_t, int64_t x_stride_b, const float* h0, const float* P,
                            const dcgru_cell_params* w, const float* h_seq, const float* ruc,
                            const float* d_hseq, const float* d_hlast, float* dx, float* dh0,
                            const dcgru_cell_grads* g, const void* gsave, size_t gsave_bytes,
                            void* workspace, size_t workspace_bytes, void* stream) {
    return encoder_layer_bwd_impl(d, batch, seq_len, x, x_stride_t, x_stride_b, h0, P, w, h_seq, ruc, d_hseq, d_hlast, nullptr,
                                  nullptr, dx, dh0, g, gsave, gsave_bytes, workspace, workspace_bytes, stream);
}

size_t dcgru_encoder_layer_bwd_sel_workspace(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len) {
    if (check_desc(d) || batch < 1 || seq_len < 1) return 0;
    return align_up(dcgru_encoder_layer_bwd_workspace(d, batch, seq_len)) + sel_dense_bytes(d, batch, seq_len);
}

int dcgru_encoder_layer_bwd_sel(const dcgru_cell_desc* d, int32_t batch, int32_t seq_len, const float* x,
                                int64_t x_stride_t, int64_t x_stride_b, const float* h0, const float* P,
                                const dcgru_cell_params* w, const float* h_seq, const float* ruc,
                                const float* d_hsel, const int32_t* sel_t, const float* d_hlast, float* dx, float* dh0,
                                const dcgru_cell_grads* g, const void* gsave, size_t gsave_bytes,
                                void* workspace, size_t workspace_bytes, void* stream) {
    if (!d_hsel) return fail("null d_hsel");
    return encoder_layer_bwd_impl(d, batch, seq_len, x, x_stride_t, x_stride_b, h0, P, w, h_seq, ruc, nullptr, d_hlast, d_hsel,
                                  sel_t, dx, dh0, g, gsave, gsave_bytes, workspace, workspace_bytes, stream);
}

// ---- fused classification head (head.cu) ----------------------------------------------------------------------------------
int dcgru_cls_head_fwd(int32_t batch, int32_t seq_len, int32_t num_nodes, int32_t hid_dim, int32_t num_classes,
                       const float* h_seq, const int32_t* sel_t, const float* drop_mask, const float* fc_w, const float* fc_b,
                       float* logits, int32_t* argmax_node, void* stream) {
    if (batch < 1 || seq_len < 1) return fail("empty batch/sequence");
    if (!cls_head_supported(num_nodes, hid_dim, num_classes))
        return fail("cls_head: unsupported shape (N=%d H=%d classes=%d)", num_nodes, hid_dim, num_classes);
    if (!h_seq || !fc_w || !fc_b || !logits || !argmax_node) return fail("null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("cls_head_fwd", launch_cls_head_fwd(batch, seq_len, num_nodes, hid_dim, num_classes, h_seq, sel_t, drop_mask, fc_w,
                                               fc_b, logits, argmax_node, st));
    return 0;
}

size_t dcgru_cls_head_bwd_workspace(int32_t batch, int32_t hid_dim, int32_t num_classes) {
    if (batch < 1 || hid_dim < 1 || num_classes < 1) return 0;
    return align_up((size_t)batch * num_classes * hid_dim * 4);
}

int dcgru_cls_head_bwd(int32_t batch, int32_t seq_len, int32_t num_nodes, int32_t hid_dim, int32_t num_classes,
                       const float* h_seq, const int32_t* sel_t, const float* drop_mask, const float* fc_w,
                       const int32_t* argmax_node, const float* d_logits, float* d_hsel, float* d_fc_w, float* d_fc_b,
                       void* workspace, size_t workspace_bytes, void* stream) {
    if (batch < 1 || seq_len < 1) return fail("empty batch/sequence");
    if (!cls_head_supported(num_nodes, hid_dim, num_classes))
        return fail("cls_head: unsupported shape (N=%d H=%d classes=%d)", num_nodes, hid_dim, num_classes);
    if (!h_seq || !fc_w || !argmax_node || !d_logits || !d_hsel || !d_fc_w || !d_fc_b || !workspace) return fail("null pointer");
    if (workspace_bytes < dcgru_cls_head_bwd_workspace(batch, hid_dim, num_classes)) return fail("workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("cls_head_bwd", launch_cls_head_bwd(batch, seq_len, num_nodes, hid_dim, num_classes, h_seq, sel_t, drop_mask, fc_w,
                                               argmax_node, d_logits, d_hsel, d_fc_w, d_fc_b, reinterpret_cast<float*>(workspace),
                                               st));
    return 0;
}

// ---- second-generation kernels: stage-level debug entry points (tests/test_gpu_g2.py) ---------------------------------
size_t dcgru_debug_bulk_dp_workspace(int32_t mode, int32_t fin, int32_t H, int32_t M) {
    return align_up(mode == 0 ? bulk_wimg_bytes(fin, M, 3 * H) : bulk_wimg_bytes(3 * H, M, fin));
}

int dcgru_debug_bulk_dp(int32_t mode, int32_t B, int32_t T, int32_t N, int32_t fin, int32_t H, int32_t M, const float* src,
                        const float* P, const float* Wg, const float* Wc, const float* bias, float* out, void* img,
                        int32_t img_cols, int32_t img_col0, void* workspace, size_t workspace_bytes, void* stream) {
    if (!src || !Wg || !Wc || !out || !workspace) return fail("null pointer");
    if (M > 1 && !P) return fail("null P");
    if (H != 64) return fail("hid_dim=%d unsupported by the 2xFP16 kernels", H);
    const int Cin = mode == 0 ? fin : 3 * H, Nout = mode == 0 ? 3 * H : fin;
    const DevInfo& di = devinfo();
    if (!bulk_dp_supported(N, Cin, M, Nout, false, di.smem)) return fail("bulk_dp: unsupported shape (N=%d Cin=%d M=%d Nout=%d)", N, Cin, M, Nout);
    if (workspace_bytes < dcgru_debug_bulk_dp_workspace(mode, fin, H, M)) return fail("workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("pack_w16", launch_pack_w16(Wg, Wc, fin, H, M, mode == 0 ? 0 : 1, Nout, g16_nq(Cin, M), workspace, st));
    LAUNCH(mode == 0 ? "xproj" : "dx16",
           launch_bulk_dp(B, T, N, Cin, M, Nout, mode == 0 ? 0 : 1, src, (long long)B * N * Cin, (long long)N * Cin, nullptr, P, workspace,
                          bias, out, (long long)B * N * Nout, (long long)N * Nout, Nout, 1.f, nullptr, img, img_cols, img_col0,
                          di.sms, di.smem, st));
    return 0;
}

// ---- fused optimiser step -------------------------------------------------------------------------------
size_t dcgru_clip_adam_workspace(size_t n) { return align_up((size_t)clip_adam_npart(n) * 8); }

int dcgru_clip_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, size_t n, const float* lr,
                         int32_t* step, float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                         float grad_scale, float* total_norm, void* workspace, size_t workspace_bytes, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !lr || !step || !workspace) return fail("null pointer");
    if (n == 0) return fail("empty parameter buffer");
    if (workspace_bytes < dcgru_clip_adam_workspace(n)) return fail("workspace too small");
    if (!(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps > 0.f)) return fail("bad Adam hyper-parameters");
    if (!(grad_scale > 0.f)) return fail("grad_scale must be positive (1 = none, 1/world = data-parallel average)");
    if (reinterpret_cast<uintptr_t>(workspace) & 7) return fail("workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    LAUNCH("clip_adam", launch_clip_adam(params, grads, exp_avg, exp_avg_sq, n, lr, step, beta1, beta2, eps, weight_decay,
                                         max_grad_norm, grad_scale, (double*)workspace, total_norm, st));
    return 0;
}

// ---- decoder ------------------------------------------------------------------------------------------
static int check_dec(const dcgru_cell_desc* d, int L, int B, int T) {
    if (check_desc(d)) return 1;
    if (L < 1 || L > DCGRU_MAX_LAYERS) return fail("num_layers=%d unsupported (1..%d)", L, DCGRU_MAX_LAYERS);
    if (B < 1 || T < 1) return fail("empty batch/sequence");
    if (T > 64) return fail("decoder seq_len=%d > 64 unsupported", T);
    return 0;
}

// ---- second-generation decoder (2xFP16 tensor-core kernels, H = 64) ---------------------------------------------------------
// The decoder is time-outer / layer-inner and autoregressive (model/model.py:182-202): nothing can be hoisted over time, so
// every (step, layer) cell is one T = 1 launch pair of the encoder's kernels -- x-part GEMM (bulk_dp.cu) + recurrent step
// (rnn_fwd.cu) -- followed per step by the projection as a bulk_dp GEMM without diffusion (M = 1).  Weight images are packed
// once per distinct cell (tied cells share one).  BPTT mirrors it with rnn_bwd.cu / the dX GEMM per cell; all dA slabs land
// in ONE operand image [tile][t*L + l], which the bulk weight-gradient kernels consume afterwards.
// DCGRU_G2_DEC=0 keeps the single-launch fp32 FMA decoder (seq_fwd.cu / seq_bwd.cu mode 1); tests compare the two.
static int nout_pad(int n) { return n <= 64 ? 64 : (n <= 128 ? 128 : 192); }
static bool g2_dec_supported(const dcgru_cell_desc* d, int L) {
    const char* e = getenv("DCGRU_G2_DEC");
    if ((e && e[0] == '0') || !g2_enabled()) return false;
    const int H = d->hid_dim, Fo = d->input_dim, M = Mof(d), N = d->num_nodes;
    if (H != 64 || Fo > 192 || Fo % 4 || L < 1) return false;
    dcgru_cell_desc d1 = *d;
    d1.input_dim = H;
    const int smem = devinfo().smem;
    if (!g2_bwd_supported(d) || !g2_bwd_supported(&d1)) return false;
    return bulk_dp_supported(N, 3 * H, M, nout_pad(Fo), true, smem) && bulk_dp_supported(N, 3 * H, M, H, true, smem) &&
           bulk_dp_supported(N, H, 1, nout_pad(Fo), false, smem) && bulk_dp_supported(N, Fo, 1, H, false, smem);
}
// distinct cells of the stack: cid[l] = first layer with the same parameters
static int dec_cell_ids(const dcgru_cell_params* w, int L, int* cid) {
    int n = 0;
    for (int l = 0; l < L; ++l) {
        cid[l] = l;
        for (int j = 0; j < l; ++j)
            if (w[j].Wg == w[l].Wg && w[j].Wc == w[l].Wc) { cid[l] = cid[j]; break; }
        if (cid[l] == l) ++n;
    }
    return n;
}
// operand images of the decoder (saved for backward): cell 0 has T slabs (slab = t), the upper cells share one image of
// T*(L-1) slabs (slab = t*(L-1) + l-1), which is the whole K range of the tied cell's weight-gradient GEMM
struct G2DecImg { int kkp0, kkp1, kxp0, kxp1; size_t bytes0, bytes1; };
static G2DecImg g2_dec_img(const dcgru_cell_desc* d, int L, int B, int T) {
    G2DecImg g;
    const int H = d->hid_dim, Fo = d->input_dim, M = Mof(d);
    g.kkp0 = g16_kkp(Fo, H, M); g.kxp0 = g16_kxp(Fo, M);
    g.kkp1 = g16_kkp(H, H, M);  g.kxp1 = g16_kxp(H, M);
    g.bytes0 = align_up(g16_image_bytes(B, T, g.kkp0));
    g.bytes1 = L > 1 ? align_up(g16_image_bytes(B, T * (L - 1), g.kkp1)) : 0;
    return g;
}
static size_t g2_dec_gsave_bytes(const dcgru_cell_desc* d, int L, int B, int T) {
    if (!g2_dec_supported(d, L) || !g2_gsave_enabled() || dw_mm16_smem_bytes() + 2048 > devinfo().smem) return 0;
    const G2DecImg g = g2_dec_img(d, L, B, T);
    return g.bytes0 + g.bytes1;
}
struct G2DecFwdWs { size_t off_wx[DCGRU_MAX_LAYERS], off_wh[DCGRU_MAX_LAYERS], off_bias[DCGRU_MAX_LAYERS], off_wp, off_xp, off_zero, off_hm, total; };
static G2DecFwdWs g2_dec_fwd_ws(const dcgru_cell_desc* d, int L, int B) {
    G2DecFwdWs w;
    const int H = d->hid_dim, Fo = d->input_dim, M = Mof(d), N = d->num_nodes;
    size_t o = 0;
    for (int l = 0; l < L; ++l) {
        const int fin = l == 0 ? Fo : H;
        w.off_wx[l] = o; o = align_up(o + bulk_wimg_bytes(fin, M, 3 * H));
        w.off_wh[l] = o; o = align_up(o + rnn_fwd_wimg_bytes(M));
        w.off_bias[l] = o; o = align_up(o + (size_t)3 * H * 4);
    }
    w.off_wp = o; o = align_up(o + bulk_wimg_bytes(H, 1, nout_pad(Fo)));
    w.off_xp = o; o = align_up(o + (size_t)B * N * 3 * H * 4);
    w.off_zero = o; o = align_up(o + (size_t)B * N * Fo * 4);
    w.off_hm = o; o = align_up(o + (size_t)B * N * H * 4);
    w.total = o;
    return w;
}

size_t dcgru_decoder_fwd_workspace(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T) {
    if (check_dec(d, L, B, T)) return 0;
    // projWT (H, FoPad) with the largest possible padding (SB = 1 -> multiples of 256)
    size_t n = align_up((size_t)d->hid_dim * ((d->input_dim + 255) / 256 * 256) * 4);
    if (g2_dec_supported(d, L)) { const size_t g = g2_dec_fwd_ws(d, L, B).total; if (g > n) n = g; }
    return n;
}

static int g2_decoder_fwd(const dcgru_cell_desc* d, int L, int B, int T, const float* targets, uint64_t teacher_mask,
                          const float* h0, const float* P, const dcgru_cell_params* w, const float* proj_w, const float* proj_b,
                          const float* drop_mask, float* out, float* h_all, float* ruc, void* gsave, void* workspace,
                          cudaStream_t st) {
    const int M = Mof(d), H = d->hid_dim, Fo = d->input_dim, N = d->num_nodes;
    const DevInfo& di = devinfo();
    const G2DecFwdWs ws = g2_dec_fwd_ws(d, L, B);
    const G2DecImg gi = g2_dec_img(d, L, B, T);
    uint8_t* img0 = reinterpret_cast<uint8_t*>(gsave);
    uint8_t* img1 = gsave ? img0 + gi.bytes0 : nullptr;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
    int cid[DCGRU_MAX_LAYERS];
    dec_cell_ids(w, L, cid);
    for (int l = 0; l < L; ++l) {
        if (cid[l] != l) continue;
        const int fin = l == 0 ? Fo : H;
        float* bias = reinterpret_cast<float*>(wsb + ws.off_bias[l]);
        CUDA_TRY(cudaMemcpyAsync(bias, w[l].bg, (size_t)2 * H * 4, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(bias + 2 * H, w[l].bc, (size_t)H * 4, cudaMemcpyDeviceToDevice, st));
        LAUNCH("pack_w16", launch_pack_w16(w[l].Wg, w[l].Wc, fin, H, M, 0, 3 * H, g16_nq(fin, M), wsb + ws.off_wx[l], st));
        LAUNCH("pack_w16", rnn_fwd_pack_weights(w[l].Wg, w[l].Wc, fin, M, wsb + ws.off_wh[l], st));
    }
    const int fop = nout_pad(Fo);
    LAUNCH("pack_w16", launch_pack_w16(proj_w, nullptr, Fo, H, 1, 6, fop, 1, wsb + ws.off_wp, st));
    float* xp = reinterpret_cast<float*>(wsb + ws.off_xp);
    float* zero = reinterpret_cast<float*>(wsb + ws.off_zero);
    float* hm = reinterpret_cast<float*>(wsb + ws.off_hm);
    CUDA_TRY(cudaMemsetAsync(zero, 0, (size_t)B * N * Fo * 4, st));                         // GO symbol (model/model.py:172-173)
    const size_t NH = (size_t)N * H, NF = (size_t)N * Fo;
    BulkExtra pex;
    pex.nout_valid = Fo;
    for (int t = 0; t < T; ++t) {
        const float* xin = zero;
        if (t > 0) xin = ((teacher_mask >> (t - 1)) & 1) ? targets + (size_t)(t - 1) * B * NF : out + (size_t)(t - 1) * B * NF;
        for (int l = 0; l < L; ++l) {
            const int fin = l == 0 ? Fo : H, c = cid[l];
            const float* src = l == 0 ? xin : h_all + ((size_t)t * L + (l - 1)) * B * NH;
            const float* hprev = t == 0 ? h0 + (size_t)l * B * NH : h_all + ((size_t)(t - 1) * L + l) * B * NH;
            float* hout = h_all + ((size_t)t * L + l) * B * NH;
            float* rucl = ruc ? ruc + ((size_t)t * L + l) * B * NH * 3 : nullptr;
            void* img = l == 0 ? img0 : img1;
            const int kkp = l == 0 ? gi.kkp0 : gi.kkp1, kxp = l == 0 ? gi.kxp0 : gi.kxp1;
            BulkExtra xex;
            xex.img_T = l == 0 ? T : T * (L - 1);
            xex.img_t0 = l == 0 ? t : t * (L - 1) + (l - 1);
            LAUNCH("xproj", launch_bulk_dp(B, 1, N, fin, M, 3 * H, 0, src, 0, (long long)N * fin, nullptr, P, wsb + ws.off_wx[c],
                                           reinterpret_cast<float*>(wsb + ws.off_bias[c]), xp, 0, (long long)N * 3 * H, 3 * H, 1.f,
                                           nullptr, img, kkp, 0, di.sms, di.smem, st, &xex));
            LAUNCH("rnn_fwd", launch_rnn_fwd(B, 1, N, fin, M, d->activation, xp, hprev, P, nullptr, nullptr, wsb + ws.off_wh[c], hout,
                                             rucl, img, kkp, kxp, st, xex.img_T, xex.img_t0));
        }
        const float* top = h_all + ((size_t)t * L + (L - 1)) * B * NH;
        if (drop_mask) {                                                                     // nn.Dropout before the projection (model/model.py:192)
            LAUNCH("ew", launch_ew_add_mul(top, nullptr, drop_mask + (size_t)t * B * NH, hm, (size_t)B * NH, st));
            top = hm;
        }
        LAUNCH("proj", launch_bulk_dp(B, 1, N, H, 1, fop, 0, top, 0, (long long)NH, nullptr, nullptr, wsb + ws.off_wp, proj_b,
                                      out + (size_t)t * B * NF, 0, (long long)NF, Fo, 1.f, nullptr, nullptr, 0, 0, di.sms, di.smem, st,
                                      &pex));
    }
    return 0;
}

// layers >= 1 all share one cell (the reference's decoder, model/model.py:126,142-143) -- or there is at most one of them
static bool dec_upper_tied(const dcgru_cell_params* w, int L) {
    for (int l = 2; l < L; ++l)
        if (w[l].Wg != w[1].Wg || w[l].Wc != w[1].Wc) return false;
    return true;
}

size_t dcgru_decoder_gsave_bytes(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T) {
    if (check_dec(d, L, B, T)) return 0;
    return g2_dec_gsave_bytes(d, L, B, T);
}

static int decoder_fwd_impl(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T, const float* targets,
                            uint64_t teacher_mask, const float* h0, const float* P, const dcgru_cell_params* w,
                            const float* proj_w, const float* proj_b, const float* drop_mask, float* out,
                            float* h_all, float* ruc, void* gsave, size_t gsave_bytes, void* workspace, size_t workspace_bytes,
                            void* stream) {
    if (check_dec(d, L, B, T)) return 1;
    if (!h0 || !w || !proj_w || !proj_b || !out || !h_all || !workspace) return fail("null pointer");
    if (teacher_mask && !targets) return fail("teacher forcing requested without targets");
    const int M = Mof(d), H = d->hid_dim, Fo = d->input_dim;
    if (M > 1 && !P) return fail("null P");
    if (dcgru_decoder_fwd_workspace(d, L, B, T) > workspace_bytes) return fail("workspace too small");
    for (int l = 0; l < L; ++l)
        if (!w[l].Wg || !w[l].bg || !w[l].Wc || !w[l].bc) return fail("null weights (layer %d)", l);
    if (g2_dec_supported(d, L) && aligned16(workspace) && aligned16(h0) && aligned16(h_all) && aligned16(out) &&
        (!targets || aligned16(targets)) && (!drop_mask || aligned16(drop_mask)) && (!ruc || aligned16(ruc))) {
        if (gsave) {
            const size_t need = g2_dec_gsave_bytes(d, L, B, T);
            if (need == 0) return fail("this configuration has no operand image: pass gsave = NULL");
            if (gsave_bytes < need) return fail("gsave too small (%zu < %zu bytes)", gsave_bytes, need);
            if (!aligned16(gsave)) return fail("gsave must be 16-byte aligned");
        }
        return g2_decoder_fwd(d, L, B, T, targets, teacher_mask, h0, P, w, proj_w, proj_b, drop_mask, out, h_all, ruc, gsave,
                              workspace, (cudaStream_t)stream);
    }
    if (gsave) return fail("gsave given but the tensor-core decoder path is not available for this call");
    int cmax = (Fo > H ? Fo : H) + H;
    FwdPlan pl;
    if (!plan_fwd(H, cmax, M, B, Fo, &pl)) return fail("no decoder tiling fits shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    float* projWT = (float*)workspace;
    LAUNCH("transpose", launch_transpose(proj_w, Fo, H, projWT, pl.FoPad, st));
    FwdParams p;
    memset(&p, 0, sizeof p);
    p.B = B; p.T = T; p.N = d->num_nodes; p.H = H; p.M = M; p.act = d->activation; p.ncell = L; p.KC = pl.KC; p.mode = 1;
    for (int l = 0; l < L; ++l) {
        if (!w[l].Wg || !w[l].bg || !w[l].Wc || !w[l].bc) return fail("null weights (layer %d)", l);
        p.cell[l] = CellW{w[l].Wg, w[l].bg, w[l].Wc, w[l].bc, l == 0 ? Fo : H};
    }
    p.P = P; p.h0 = h0; p.hseq = h_all; p.ruc = ruc; p.targets = targets; p.teacher_mask = teacher_mask;
    p.projWT = projWT; p.projb = proj_b; p.dropmask = drop_mask; p.out = out; p.Fo = Fo; p.FoPad = pl.FoPad;
    LAUNCH("seq_fwd", launch_seq_fwd(p, pl.SB, pl.smem, st));
    return 0;
}

int dcgru_decoder_fwd(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T, const float* targets,
                      uint64_t teacher_mask, const float* h0, const float* P, const dcgru_cell_params* w,
                      const float* proj_w, const float* proj_b, const float* drop_mask, float* out,
                      float* h_all, float* ruc, void* workspace, size_t workspace_bytes, void* stream) {
    return decoder_fwd_impl(d, L, B, T, targets, teacher_mask, h0, P, w, proj_w, proj_b, drop_mask, out, h_all, ruc, nullptr, 0,
                            workspace, workspace_bytes, stream);
}
int dcgru_decoder_fwd_saved(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T, const float* targets,
                            uint64_t teacher_mask, const float* h0, const float* P, const dcgru_cell_params* w,
                            const float* proj_w, const float* proj_b, const float* drop_mask, float* out,
                            float* h_all, float* ruc, void* gsave, size_t gsave_bytes, void* workspace, size_t workspace_bytes,
                            void* stream) {
    return decoder_fwd_impl(d, L, B, T, targets, teacher_mask, h0, P, w, proj_w, proj_b, drop_mask, out, h_all, ruc, gsave,
                            gsave_bytes, workspace, workspace_bytes, stream);
}

struct DecWs {
    float *WgT[DCGRU_MAX_LAYERS], *WcT[DCGRU_MAX_LAYERS];
    float *ptbuf;
    float *dA, *dY, *scratch, *part0, *partb0, *part1, *partb1, *partp, *partpb;
    int ns0, nj0, ns1, nj1, nsp, njp;
    bool tc0, tc1;
    float* g2;                 // second-generation BPTT scratch (g2_dec_bwd_ws) or nullptr
    size_t bytes;
};
struct G2DecBwdWs { size_t off_wb[DCGRU_MAX_LAYERS], off_wdx[DCGRU_MAX_LAYERS], off_wpb, off_img0, off_img1, off_scale, off_above, off_dxin,
                           off_part, off_cs, total; };
static G2DecBwdWs g2_dec_bwd_ws(const dcgru_cell_desc* d, int L, int B, int T) {
    G2DecBwdWs w;
    const int H = d->hid_dim, Fo = d->input_dim, M = Mof(d), N = d->num_nodes;
    size_t o = 0;
    for (int l = 0; l < L; ++l) {
        const int fin = l == 0 ? Fo : H;
        w.off_wb[l] = o; o = align_up(o + rnn_bwd_wimg_bytes(M));
        w.off_wdx[l] = o; o = align_up(o + bulk_wimg_bytes(3 * H, M, nout_pad(fin)));
    }
    w.off_wpb = o; o = align_up(o + bulk_wimg_bytes(Fo, 1, H));
    w.off_img0 = o; o = align_up(o + g16_image_bytes(B, T, 3 * H));                       // dA image of cell 0, slab = t
    w.off_img1 = o; o = align_up(o + g16_image_bytes(B, T * (L > 1 ? L - 1 : 0), 3 * H));   // upper cells, slab = t*(L-1) + l-1
    w.off_part = o; o = align_up(o + dw_mm16_part_floats(devinfo().sms) * 4);
    w.off_cs = o; o = align_up(o + colsum16_part_floats(H) * 4);
    w.off_scale = o; o = align_up(o + 256);
    w.off_above = o; o = align_up(o + (size_t)B * N * H * 4);
    w.off_dxin = o; o = align_up(o + (size_t)B * N * Fo * 4);
    w.total = o;
    return w;
}
static void dec_bwd_ws(const dcgru_cell_desc* d, int L, int B, int T, void* ws, DecWs* o) {
    const int H = d->hid_dim, M = Mof(d), Fo = d->input_dim, N = d->num_nodes;
    const int CM0 = (Fo + H) * M, CM1 = 2 * H * M;
    DwJob tmp[DW_MAXJOBS];
    { CellDwPlan a = plan_cell_dw(Fo, H, M, B, T, tmp); o->nj0 = a.njobs; o->ns0 = a.nsplit; o->tc0 = a.tc; }
    { CellDwPlan a = plan_cell_dw(H, H, M, B, T, tmp);  o->nj1 = a.njobs; o->ns1 = a.nsplit; o->tc1 = a.tc; }
    o->njp = build_proj_jobs(Fo, H, tmp);     o->nsp = dw_nsplit(B, o->njp);
    Carver c(ws, 0);
    for (int l = 0; l < L; ++l) {
        int CM = l == 0 ? CM0 : CM1;
        o->WgT[l] = c.take((size_t)2 * H * CM);
        o->WcT[l] = c.take((size_t)H * CM);
    }
    o->dA = c.take((size_t)T * L * B * N * 3 * H);
    o->dY = c.take((size_t)T * B * N * Fo);
    o->scratch = c.take((size_t)B * N * (Fo > H ? Fo : H));
    o->part0 = c.take((size_t)o->ns0 * CM0 * 3 * H);
    o->partb0 = c.take((size_t)o->ns0 * 3 * H);
    o->part1 = c.take((size_t)(L > 1 ? L - 1 : 0) * o->ns1 * CM1 * 3 * H);
    o->partb1 = c.take((size_t)(L > 1 ? L - 1 : 0) * o->ns1 * 3 * H);
    o->partp = c.take((size_t)o->nsp * Fo * H);
    o->partpb = c.take((size_t)o->nsp * Fo);
    o->ptbuf = c.take(dw_tc_pt_floats(B, M));
    o->g2 = nullptr;
    if (g2_dec_supported(d, L)) o->g2 = c.take(g2_dec_bwd_ws(d, L, B, T).total / 4);
    o->bytes = c.off;
}

size_t dcgru_decoder_bwd_workspace(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T) {
    if (check_dec(d, L, B, T)) return 0;
    DecWs o;
    dec_bwd_ws(d, L, B, T, nullptr, &o);
    return o.bytes;
}

static int decoder_bwd_impl(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T, const float* targets,
                            uint64_t teacher_mask, const float* h0, const float* P, const dcgru_cell_params* w,
                            const float* proj_w, const float* drop_mask, const float* out, const float* h_all,
                            const float* ruc, const float* d_out, float* dh0, const dcgru_cell_grads* g,
                            float* dproj_w, float* dproj_b, const void* gsave, size_t gsave_bytes, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (check_dec(d, L, B, T)) return 1;
    if (!h0 || !w || !proj_w || !out || !h_all || !ruc || !d_out || !dh0 || !g || !dproj_w || !dproj_b || !workspace)
        return fail("null pointer");
    if (teacher_mask && !targets) return fail("teacher forcing requested without targets");
    const int M = Mof(d), H = d->hid_dim, Fo = d->input_dim, N = d->num_nodes;
    if (M > 1 && !P) return fail("null P");
    // weight tying (model/model.py:126,142-143): layers >= 1 either all share one cell or none do
    bool tied = L > 2;
    for (int l = 2; l < L; ++l) tied = tied && (w[l].Wg == w[1].Wg);
    for (int l = 2; l < L; ++l)
        if (!tied && w[l].Wg == w[1].Wg) return fail("partially tied decoder cells are unsupported");
    if (tied)
        for (int l = 2; l < L; ++l)
            if (g[l].dWg != g[1].dWg) return fail("gradient buffers of tied cells must alias");
    DecWs o;
    dec_bwd_ws(d, L, B, T, workspace, &o);
    if (o.bytes > workspace_bytes) return fail("workspace too small");
    if (o.nj0 > DW_MAXJOBS || o.nj1 > DW_MAXJOBS || o.njp > DW_MAXJOBS) return fail("too many weight-gradient jobs");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t NH = (size_t)N * H;
    const bool tied_or_single = dec_upper_tied(w, L);
    bool cells_done = false;
    const bool use_g2 = o.g2 && aligned16(workspace) && aligned16(h0) && aligned16(h_all) && aligned16(ruc) && aligned16(d_out) &&
                        aligned16(dh0) && (!drop_mask || aligned16(drop_mask));
    BwdPlan pl;
    if (!use_g2 && !plan_bwd(H, M, B, Fo, &pl)) return fail("no decoder backward tiling fits shared memory");
    BwdParams p;
    memset(&p, 0, sizeof p);
    p.B = B; p.T = T; p.N = N; p.H = H; p.M = M; p.act = d->activation; p.ncell = L; p.mode = 1;
    if (use_g2) {
        // second-generation BPTT: one rnn_bwd (T = 1) + dX GEMM per (step, layer) cell, newest step first
        const G2DecBwdWs ws = g2_dec_bwd_ws(d, L, B, T);
        const DevInfo& di = devinfo();
        uint8_t* wsb = reinterpret_cast<uint8_t*>(o.g2);
        int cid[DCGRU_MAX_LAYERS];
        dec_cell_ids(w, L, cid);
        for (int l = 0; l < L; ++l) {
            if (cid[l] != l) continue;
            const int fin = l == 0 ? Fo : H;
            LAUNCH("pack_w16", rnn_bwd_pack_weights(w[l].Wg, w[l].Wc, fin, M, wsb + ws.off_wb[l], st));
            LAUNCH("pack_w16", launch_pack_w16(w[l].Wg, w[l].Wc, fin, H, M, 1, nout_pad(fin), g16_nq(3 * H, M), wsb + ws.off_wdx[l], st));
        }
        LAUNCH("pack_w16", launch_pack_w16(proj_w, nullptr, Fo, H, 1, 7, H, g16_nq(Fo, 1), wsb + ws.off_wpb, st));
        float* scale = reinterpret_cast<float*>(wsb + ws.off_scale);
        float* above = reinterpret_cast<float*>(wsb + ws.off_above);
        float* dxin = reinterpret_cast<float*>(wsb + ws.off_dxin);
        uint8_t* dimg0 = wsb + ws.off_img0;
        uint8_t* dimg1 = wsb + ws.off_img1;
        const size_t NF = (size_t)N * Fo;
        if (gsave) {
            const size_t need = g2_dec_gsave_bytes(d, L, B, T);
            if (need == 0) return fail("this configuration has no operand image: pass gsave = NULL");
            if (gsave_bytes < need) return fail("gsave too small (%zu < %zu bytes)", gsave_bytes, need);
        }
        // one power-of-two scale for every fp16 gradient operand of the launch, from the upstream gradient's magnitude
        LAUNCH("grad_scale", launch_grad_scale(d_out, (size_t)T * B * NF, nullptr, 0, nullptr, 0, reinterpret_cast<unsigned*>(scale + 8),
                                               scale, st));
        CUDA_TRY(cudaMemsetAsync(dh0, 0, (size_t)L * B * NH * 4, st));                      // carried dh_{t-1} of every layer; dh0 at the end
        for (int t = T - 1; t >= 0; --t) {
            // total gradient of out_t: upstream + what step t+1's cell 0 sent back, unless that step was teacher-forced
            const bool fed_back = t + 1 < T && !((teacher_mask >> t) & 1);
            float* dYt = o.dY + (size_t)t * B * NF;
            LAUNCH("ew", launch_ew_add_mul(d_out + (size_t)t * B * NF, fed_back ? dxin : nullptr, nullptr, dYt, (size_t)B * NF, st));
            BulkExtra pex;
            pex.in_scale_ptr = scale;
            LAUNCH("dproj", launch_bulk_dp(B, 1, N, Fo, 1, H, 0, dYt, 0, (long long)NF, nullptr, nullptr, wsb + ws.off_wpb, nullptr,
                                           above, 0, (long long)NH, H, 1.f, scale, nullptr, 0, 0, di.sms, di.smem, st, &pex));
            if (drop_mask) LAUNCH("ew", launch_ew_add_mul(above, nullptr, drop_mask + (size_t)t * B * NH, above, (size_t)B * NH, st));
            for (int l = L - 1; l >= 0; --l) {
                const int fin = l == 0 ? Fo : H, c = cid[l];
                const float* hprev = t == 0 ? h0 + (size_t)l * B * NH : h_all + ((size_t)(t - 1) * L + l) * B * NH;
                float* carry = dh0 + (size_t)l * B * NH;
                void* img = l == 0 ? dimg0 : dimg1;
                const int img_T = l == 0 ? T : T * (L - 1), img_t0 = l == 0 ? t : t * (L - 1) + (l - 1);
                LAUNCH("rnn_bwd", launch_rnn_bwd(B, 1, N, fin, M, d->activation, hprev, hprev, ruc + ((size_t)t * L + l) * B * NH * 3, P,
                                                 nullptr, nullptr, above, carry, nullptr, nullptr, wsb + ws.off_wb[c], scale, carry, img,
                                                 st, img_T, img_t0));
                // input gradient: to the layer below, or (cell 0) back to out_{t-1} when that fed this step
                const bool need_dx = l > 0 || (t > 0 && !((teacher_mask >> (t - 1)) & 1));
                if (need_dx) {
                    BulkExtra xex;
                    xex.nout_valid = fin; xex.src_T = img_T; xex.src_t0 = img_t0;
                    LAUNCH("dx16", launch_bulk_dp(B, 1, N, 3 * H, M, nout_pad(fin), 1, nullptr, 0, 0, img, P, wsb + ws.off_wdx[c], nullptr,
                                                  l > 0 ? above : dxin, 0, (long long)N * fin, fin, 1.f, scale, nullptr, 0, 0, di.sms,
                                                  di.smem, st, &xex));
                }
            }
        }
        if (gsave && tied_or_single) {
            // weight gradients of the cells: GEMMs over the saved operand images (forward: G, here: dA), bias: column sums
            const G2DecImg gi = g2_dec_img(d, L, B, T);
            const uint8_t* g0 = reinterpret_cast<const uint8_t*>(gsave);
            float* part = reinterpret_cast<float*>(wsb + ws.off_part);
            float* cs = reinterpret_cast<float*>(wsb + ws.off_cs);
            bool fuse_db = true;                                        // db inside the weight-gradient GEMM (see the encoder path)
            { const char* e = getenv("DCGRU_FUSE_DB"); if (e && e[0] == '0') fuse_db = false; }
            LAUNCH("dw_mm16", launch_dw_mm16(Fo, H, M, B, T, g0, dimg0, part, scale, di.sms, g[0].dWg, g[0].dWc, st, fuse_db ? cs : nullptr,
                                             g[0].dbg, g[0].dbc));
            if (!fuse_db) LAUNCH("colsum16", launch_colsum16(dimg0, B, T, H, cs, scale, g[0].dbg, g[0].dbc, st));
            if (L > 1) {
                LAUNCH("dw_mm16", launch_dw_mm16(H, H, M, B, T * (L - 1), g0 + gi.bytes0, dimg1, part, scale, di.sms, g[1].dWg, g[1].dWc, st,
                                                 fuse_db ? cs : nullptr, g[1].dbg, g[1].dbc));
                if (!fuse_db) LAUNCH("colsum16", launch_colsum16(dimg1, B, T * (L - 1), H, cs, scale, g[1].dbg, g[1].dbc, st));
            }
            cells_done = true;
        } else {
            // row-major fp32 dA (T,L,B,N,3H) for the recompute weight-gradient kernels below
            LAUNCH("img_to_rows", launch_img_to_rows(dimg0, B, T, N, 3 * H, scale, o.dA, 1, L, 0, st));
            if (L > 1) LAUNCH("img_to_rows", launch_img_to_rows(dimg1, B, T * (L - 1), N, 3 * H, scale, o.dA, L - 1, L, 1, st));
        }
    } else if (gsave) {
        return fail("gsave given but the tensor-core decoder path is not available for this call");
    }
    for (int l = 0; l < L && !use_g2; ++l) {
        int fin = l == 0 ? Fo : H, CM = (fin + H) * M;
        if (l <= 1 || !tied) {
            LAUNCH("transpose", launch_transpose(w[l].Wg, CM, 2 * H, o.WgT[l], CM, st));
            LAUNCH("transpose", launch_transpose(w[l].Wc, CM, H, o.WcT[l], CM, st));
            p.cell[l] = CellWT{o.WgT[l], o.WcT[l], fin};
        } else {
            p.cell[l] = p.cell[1];
        }
    }
    p.P = P; p.h0 = h0; p.hseq = h_all; p.ruc = ruc; p.dh0 = dh0; p.dA = o.dA;
    p.d_out = d_out; p.proj_w = proj_w; p.dropmask = drop_mask; p.teacher_mask = teacher_mask;
    p.dY = o.dY; p.scratch = o.scratch; p.Fo = Fo;
    if (!use_g2) {
        CUDA_TRY(cudaMemsetAsync(dh0, 0, (size_t)L * B * NH * 4, st));
        LAUNCH("seq_bwd", launch_seq_bwd(p, pl.SB, pl.smem, st));
    }
    // ---- bulk gradients ----------------------------------------------------------------------------
    DwParams q;
    memset(&q, 0, sizeof q);
    q.B = B; q.T = T; q.N = N; q.H = H; q.M = M; q.mode = 1; q.ncell = L; q.Fo = Fo;
    q.P = P; q.h0 = h0; q.hseq = h_all; q.ruc = ruc; q.dA = o.dA; q.targets = targets; q.out = out;
    q.teacher_mask = teacher_mask; q.dY = o.dY; q.dropmask = drop_mask;
    // cell 0
    if (!cells_done) {
    if (o.tc0 || o.tc1) LAUNCH("make_pt", launch_make_pt(P, B, M, N, o.ptbuf, st));
    q.layer = 0; q.fin = Fo; q.nsplit = o.ns0; q.part = o.part0; q.partb = o.partb0;
    q.dY = o.tc0 ? o.ptbuf : o.dY;
    plan_cell_dw(Fo, H, M, B, T, q.jobs);
    if (o.tc0) LAUNCH("dw_tc", launch_dw_tc(q, o.nj0, otile(3 * H), st));
    else LAUNCH("dw", launch_dw(q, o.nj0, otile(3 * H), st));
    LAUNCH("reduce", launch_reduce_cell(o.part0, o.partb0, o.ns0, (Fo + H) * M, H, g[0].dWg, g[0].dbg, g[0].dWc, g[0].dbc, st));
    // cells >= 1
    plan_cell_dw(H, H, M, B, T, q.jobs);
    const size_t psz = (size_t)2 * H * M * 3 * H;
    for (int l = 1; l < L; ++l) {
        q.layer = l; q.fin = H; q.nsplit = o.ns1; q.dY = o.tc1 ? o.ptbuf : o.dY;
        q.part = o.part1 + (size_t)(l - 1) * o.ns1 * psz;
        q.partb = o.partb1 + (size_t)(l - 1) * o.ns1 * 3 * H;
        if (o.tc1) LAUNCH("dw_tc", launch_dw_tc(q, o.nj1, otile(3 * H), st));
        else LAUNCH("dw", launch_dw(q, o.nj1, otile(3 * H), st));
        if (!tied)
            LAUNCH("reduce", launch_reduce_cell(q.part, q.partb, o.ns1, 2 * H * M, H, g[l].dWg, g[l].dbg, g[l].dWc, g[l].dbc, st));
    }
    if (tied)
        LAUNCH("reduce", launch_reduce_cell(o.part1, o.partb1, (L - 1) * o.ns1, 2 * H * M, H, g[1].dWg, g[1].dbg, g[1].dWc,
                                    g[1].dbc, st));
    }
    // Linear
    q.layer = L - 1; q.fin = Fo; q.nsplit = o.nsp; q.part = o.partp; q.partb = o.partpb; q.dY = o.dY;
    build_proj_jobs(Fo, H, q.jobs);
    LAUNCH("dw", launch_dw(q, o.njp, otile(H), st));
    LAUNCH("reduce", launch_reduce_flat(o.partp, o.nsp, (size_t)Fo * H, dproj_w, st));
    LAUNCH("reduce", launch_reduce_flat(o.partpb, o.nsp, (size_t)Fo, dproj_b, st));
    return 0;
}

int dcgru_decoder_bwd(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T, const float* targets,
                      uint64_t teacher_mask, const float* h0, const float* P, const dcgru_cell_params* w,
                      const float* proj_w, const float* drop_mask, const float* out, const float* h_all,
                      const float* ruc, const float* d_out, float* dh0, const dcgru_cell_grads* g,
                      float* dproj_w, float* dproj_b, void* workspace, size_t workspace_bytes, void* stream) {
    return decoder_bwd_impl(d, L, B, T, targets, teacher_mask, h0, P, w, proj_w, drop_mask, out, h_all, ruc, d_out, dh0, g, dproj_w,
                            dproj_b, nullptr, 0, workspace, workspace_bytes, stream);
}
int dcgru_decoder_bwd_saved(const dcgru_cell_desc* d, int32_t L, int32_t B, int32_t T, const float* targets,
                            uint64_t teacher_mask, const float* h0, const float* P, const dcgru_cell_params* w,
                            const float* proj_w, const float* drop_mask, const float* out, const float* h_all,
                            const float* ruc, const float* d_out, float* dh0, const dcgru_cell_grads* g,
                            float* dproj_w, float* dproj_b, const void* gsave, size_t gsave_bytes, void* workspace,
                            size_t workspace_bytes, void* stream) {
    return decoder_bwd_impl(d, L, B, T, targets, teacher_mask, h0, P, w, proj_w, drop_mask, out, h_all, ruc, d_out, dh0, g, dproj_w,
                            dproj_b, gsave, gsave_bytes, workspace, workspace_bytes, stream);
}

}  // extern "C"
