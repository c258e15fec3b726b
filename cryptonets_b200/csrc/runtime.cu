// Host runtime of libcnhe (see runtime.h).  Product code: builds every table with hostmath.h, never touches oracle/.
#include "runtime.h"
#include <chrono>
#include <cmath>
#include <cstdio>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <sys/random.h>

#include "hostmath.h"

namespace cnhe {

// 256-bit ChaCha20 key straight from the kernel CSPRNG (getrandom(2)); no fallback: a context without entropy must not encrypt
void rng_from_os(RngKey &rk) {
    memset(&rk, 0, sizeof(rk));
    unsigned char *p = reinterpret_cast<unsigned char *>(rk.key);
    size_t got = 0;
    while (got < sizeof(rk.key)) {
        const ssize_t r = getrandom(p + got, sizeof(rk.key) - got, 0);
        if (r <= 0) throw Error(-2, "getrandom() failed: no OS entropy for key generation / encryption");
        got += (size_t)r;
    }
    rk.secure = 1;
}
const char *op_kind_name(int kind) {
    static const char *names[] = {"Encryption", "Decryption", "Multiplication", "Relinarization", "PlainMultiplication", "ScalarMultiplication",
                                  "Addition", "PlainAddition", "Subtraction", "PlainSubtraction", "Rotation", "ColumnRotation", "AddMany",
                                  "AddManyItemCount"};
    return kind >= 0 && kind < Context::OP_COUNT ? names[kind] : "?";
}
void Context::note(OpKind kind, int channel, int n, const u64 *first_out, const u64 *in0, const u64 *in1, double aux) {
    op_count[kind] += (uint64_t)n;
    if (!trace_noise || kind == OP_ADD_MANY_ITEMS) return;
    int budget = -1;
    const int b0 = in0 ? known_budget(in0) : -1, b1 = in1 ? known_budget(in1) : -1; // before the output (possibly in place) is re-measured
    if (first_out && channel >= 0 && channel < P && ch[channel].have_sk) {
        budget = op_noise_budget(*this, channel, first_out);
        // the call wrote n ciphertexts starting here: whatever was recorded for these addresses (recycled blocks) is stale now
        budget_of.erase(budget_of.lower_bound(first_out), budget_of.lower_bound(first_out + (size_t)n * ct_words()));
        budget_of[first_out] = budget;
    }
    trace.push_back({(int)kind, channel, n, budget, b0, b1, (int)std::lround(aux * 1000.0), 0});
}

void cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw Error(-2, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

// CNHE_TRACE_SLOW=<ms>: report host-side calls that take longer than that (diagnosing launch-path stalls)
static double trace_slow_ms() {
    static const double v = getenv("CNHE_TRACE_SLOW") ? atof(getenv("CNHE_TRACE_SLOW")) : 0.0;
    return v;
}
void Context::trace_gap(const char *what) {
    static thread_local std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
    const auto now = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - last).count();
    if (ms > trace_ms && ms < 1000.0) fprintf(stderr, "[cnhe] %.2f ms of host time before/in launch of %s\n", ms, what);
    last = now;
}
DevBuf::DevBuf(size_t w, cudaStream_t s) : words(w), stream(s) {
    if (!w) return;
    if (trace_slow_ms() > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        CNHE_CUDA(cudaMallocAsync((void **)&p, w * sizeof(u64), s));
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > trace_slow_ms()) fprintf(stderr, "[cnhe] slow cudaMallocAsync: %.2f ms for %.1f MB\n", ms, w * 8.0 / 1e6);
        return;
    }
    CNHE_CUDA(cudaMallocAsync((void **)&p, w * sizeof(u64), s));
}
constexpr size_t RECYCLE_MIN_WORDS = (size_t)1 << 19;        // 4 MB
constexpr size_t RECYCLE_CAP_WORDS = (size_t)40 << 27;       // 40 GB parked at most
DevBuf::DevBuf(size_t w, cudaStream_t s, Context *ctx) : words(w), stream(s), owner(ctx) {
    if (!w) return;
    if (ctx && w >= RECYCLE_MIN_WORDS) {
        size_t got = 0;
        p = ctx->take_recycled(w, s, got);
        if (p) { words = got; return; }
    }
    DevBuf fresh(w, s); // traced allocation path
    p = fresh.p;
    fresh.p = nullptr;
}
u64 *Context::take_recycled(size_t words, cudaStream_t s, size_t &got_words) {
    int best = -1;
    for (int i = 0; i < (int)recycle.size(); i++) {
        const Recycled &r = recycle[i];
        if (r.stream != s || r.words < words || r.words > words + words / 2) continue;
        if (best < 0 || r.words < recycle[best].words) best = i;
    }
    if (best < 0) return nullptr;
    u64 *p = recycle[best].p;
    got_words = recycle[best].words;
    recycle_words -= got_words;
    recycle.erase(recycle.begin() + best);
    return p;
}
bool Context::give_recycled(u64 *p, size_t words, cudaStream_t s) {
    if (!recycle_on || words < RECYCLE_MIN_WORDS || recycle_words + words > RECYCLE_CAP_WORDS || recycle.size() >= 256) return false;
    recycle.push_back({p, words, s});
    recycle_words += words;
    return true;
}
void Context::drop_recycled() {
    for (const Recycled &r : recycle) cudaFreeAsync(r.p, r.stream);
    recycle.clear();
    recycle_words = 0;
}
BufRef Context::alloc_upload(size_t words, cudaStream_t release_stream) {
    int pick = -1;
    for (int pass = 0; pass < 2 && pick < 0; pass++) // first a free slot whose last users are already done, else the longest-released one
        for (int i = 0; i < (int)upload_slots.size(); i++) {
            UploadSlot &u = upload_slots[i];
            if (u.busy || u.words < words || u.words > words + words / 2) continue;
            if (pass == 0 && cudaEventQuery(u.released) != cudaSuccess) continue;
            if (pick < 0 || u.stamp < upload_slots[pick].stamp) pick = i;
        }
    if (pick >= 0 && cudaEventQuery(upload_slots[pick].released) != cudaSuccess) { // grow to three slots of this size (enough for a
        int same = 0;                                                               // one-deep pipeline), then wait for the oldest
        for (const UploadSlot &u : upload_slots) same += u.words >= words && u.words <= words + words / 2;
        if (same < 3 * (int)streams.size()) pick = -1;
    }
    (void)cudaGetLastError(); // cudaEventQuery's cudaErrorNotReady is not an error
    if (pick < 0) { // first batch of this size: create the whole rotation at once (cudaMalloc of 0.5 GB costs ~50 ms; pay it up front)
        int same = 0;
        for (const UploadSlot &u : upload_slots) same += u.words >= words && u.words <= words + words / 2;
        const int want = same == 0 ? 3 * (int)streams.size() : same + 1;
        for (; same < want; same++) {
            UploadSlot u;
            u.words = words;
            u.busy = false;
            u.stamp = 0;
            CNHE_CUDA(cudaMalloc((void **)&u.p, words * sizeof(u64)));
            CNHE_CUDA(cudaEventCreateWithFlags(&u.released, cudaEventDisableTiming));
            upload_slots.push_back(u);
            if (pick < 0) pick = (int)upload_slots.size() - 1;
        }
    }
    UploadSlot &u = upload_slots[pick];
    u.busy = true;
    CNHE_CUDA(cudaStreamWaitEvent(copy_stream, u.released, 0)); // a never-recorded event is complete
    BufRef b = std::make_shared<DevBuf>(0, release_stream);
    b->p = u.p;
    b->words = u.words;
    b->owner = this;
    b->upload_slot = pick;
    return b;
}
void Context::release_upload(int slot, cudaStream_t s) {
    UploadSlot &u = upload_slots[slot];
    cudaEventRecord(u.released, s);
    u.busy = false;
    u.stamp = ++upload_stamp;
}
DevBuf::~DevBuf() {
    if (!p) return;
    if (upload_slot >= 0) { owner->release_upload(upload_slot, stream); return; }
    if (owner && owner->give_recycled(p, words, stream)) return;
    cudaFreeAsync(p, stream);
}

static std::vector<BufRef> &temps_of(Context &c) { return c.temps; } // per context, guarded by the context mutex

u64 *Context::ws_alloc(size_t words) {
    BufRef b = alloc(words ? words : 1);
    temps_of(*this).push_back(b);
    ws_used += words;
    return b->p;
}
void Context::sync() {
    for (cudaStream_t s : streams) CNHE_CUDA(cudaStreamSynchronize(s));
}
void Context::join_streams() {
    for (size_t i = 1; i < streams.size(); i++) {
        CNHE_CUDA(cudaEventRecord(ev_join, streams[i]));
        CNHE_CUDA(cudaStreamWaitEvent(streams[0], ev_join, 0));
    }
}
void Context::fork_streams() {
    if (streams.size() < 2) return;
    CNHE_CUDA(cudaEventRecord(ev_join, streams[0]));
    for (size_t i = 1; i < streams.size(); i++) CNHE_CUDA(cudaStreamWaitEvent(streams[i], ev_join, 0));
}
void Context::h2d(void *dst, const void *src, size_t bytes) {
    if (!bytes) return;
    if (!stage_buf) {
        stage_size = 64u << 20;
        CNHE_CUDA(cudaHostAlloc((void **)&stage_buf, stage_size, cudaHostAllocDefault));
        stage_ev.resize((size_t)STAGE_PARTS * streams.size());
        for (cudaEvent_t &e : stage_ev) CNHE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        stage_ev_set.assign(STAGE_PARTS, 0);
    }
    const size_t part_size = stage_size / STAGE_PARTS;
    if (bytes > part_size) { // large: plain (staged by the driver) copy
        CNHE_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
        return;
    }
    const size_t need = (bytes + 255) & ~(size_t)255;
    size_t part = stage_off / part_size;
    if (stage_off + need > (part + 1) * part_size || stage_off + need > stage_size) { // move on to the next part
        // what was staged in the part being left is consumed once every channel stream has passed this point
        for (size_t s = 0; s < streams.size(); s++) CNHE_CUDA(cudaEventRecord(stage_ev[part * streams.size() + s], streams[s]));
        stage_ev_set[part] = 1;
        part = (part + 1) % STAGE_PARTS;
        if (stage_ev_set[part]) // its previous contents: staged STAGE_PARTS - 1 parts ago, long consumed in the steady state
            for (size_t s = 0; s < streams.size(); s++) CNHE_CUDA(cudaEventSynchronize(stage_ev[part * streams.size() + s]));
        stage_off = part * part_size;
    }
    memcpy(stage_buf + stage_off, src, bytes);
    CNHE_CUDA(cudaMemcpyAsync(dst, stage_buf + stage_off, bytes, cudaMemcpyHostToDevice, stream));
    stage_off += need;
}
void Context::prof_begin(int family, double bytes) {
    if (!prof) return;
    ProfRec r;
    r.family = family;
    r.bytes = bytes;
    for (cudaEvent_t *e : {&r.e0, &r.e1}) {
        if (!prof_pool.empty()) { *e = prof_pool.back(); prof_pool.pop_back(); }
        else CNHE_CUDA(cudaEventCreate(e));
    }
    CNHE_CUDA(cudaEventRecord(r.e0, stream));
    prof_recs.push_back(r);
}
void Context::prof_end() {
    if (!prof) return;
    CNHE_CUDA(cudaEventRecord(prof_recs.back().e1, stream));
    if (prof_recs.size() >= 4096) prof_flush();
}
void Context::prof_flush() {
    if (prof_recs.empty()) return;
    sync();
    for (auto &r : prof_recs) {
        float ms = 0;
        CNHE_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
        prof_ms[r.family] += ms;
        prof_bytes[r.family] += r.bytes;
        prof_n[r.family]++;
        prof_pool.push_back(r.e0);
        prof_pool.push_back(r.e1);
    }
    prof_recs.clear();
}
struct ProfScope {
    Context &c;
    ProfScope(Context &ctx, int family, double bytes) : c(ctx) { c.prof_begin(family, bytes); }
    ~ProfScope() { c.prof_end(); }
};
#define PROF(family, bytes) ProfScope prof_scope_##__LINE__(c, family, bytes)
Context::~Context() {
    cudaSetDevice(device);
    for (cudaStream_t s : streams) cudaStreamSynchronize(s);
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    recycle_on = false; // buffers released from here on go straight back to the driver
    temps.clear();
    ch.clear();
    drop_recycled();
    if (d_bc) cudaFree(d_bc);
    if (d_bf) cudaFree(d_bf);
    if (d_tabs) cudaFree(d_tabs);
    if (d_table_mem) cudaFree(d_table_mem);
    if (d_index_map) cudaFree(d_index_map);
    for (auto &r : prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : prof_pool) cudaEventDestroy(e);
    if (stage_buf) cudaFreeHost(stage_buf);
    for (cudaEvent_t e : stage_ev) cudaEventDestroy(e);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (ev_join) cudaEventDestroy(ev_join);
    for (cudaStream_t s : streams) cudaStreamDestroy(s);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (UploadSlot &u : upload_slots) { cudaFree(u.p); cudaEventDestroy(u.released); }
    if (ev_copy) cudaEventDestroy(ev_copy);
    for (cudaEvent_t e : ev_export) cudaEventDestroy(e);
}
// free the temporaries of the previous operation (stream ordered, so kernels still in flight keep their memory)
void ws_release_all(Context &c) {
    temps_of(c).clear();
    c.ws_used = 0;
}
WsScope::WsScope(Context &ctx) : c(ctx), mark(temps_of(ctx).size()) {}
WsScope::~WsScope() { temps_of(c).resize(mark); }

std::vector<u64> default_coeff_modulus(uint32_t N) {
    // DefaultParams.CoeffModulus128(N) of SEAL 3.2 ("HE Wrapper/AtomicSealBfvVector.cs:146"); values = the largest primes
    // congruent to 1 mod 2N with bit sizes 54 | 36,36,37 | 43,43,44,44,44 | 48x3,49x6 (checked by tests/test_tables.py)
    switch (N) {
    case 2048: return {0x3fffffff000001ULL};
    case 4096: return {0xffffee001ULL, 0xffffc4001ULL, 0x1ffffe0001ULL};
    case 8192: return {0x7fffffd8001ULL, 0x7fffffc8001ULL, 0xfffffffc001ULL, 0xffffff6c001ULL, 0xfffffebc001ULL};
    case 16384:
        return {0xfffffffd8001ULL,  0xfffffffa0001ULL,  0xfffffff00001ULL,  0x1fffffff68001ULL, 0x1fffffff50001ULL,
                0x1ffffffee8001ULL, 0x1ffffffea0001ULL, 0x1ffffffe88001ULL, 0x1ffffffe48001ULL};
    default: return {};
    }
}

static DMod make_dmod(u64 p) {
    DMod m;
    m.p = p;
    hm::barrett_ratio(p, m.r0, m.r1);
    return m;
}
// Decide whether the FP64 butterfly path is exact for this modulus and schedule its re-centring points.
// Magnitudes are tracked in units of p; every operand of a modular product must stay below 2^52 (limit L = 2^52/p, 10% margin).
static void fp_schedule(NttTab &tb, int logN, bool force_int) {
    tb.fp_ok = 0;
    tb.fwd_recenter = tb.inv_recenter = 0;
    const u64 p = tb.mod.p;
    if (force_int || hm::bit_length(p) > 49) return;
    const double L = 0.9 * 4503599627370496.0 / (double)p;
    auto c = [&](double a) { return 0.5 + 0.75 * a / (L / 0.9) + 1e-6; }; // bound of |a*w mod p| for |a| <= a*p, |w| <= p/2
    // forward schedule of a pass list; returns false when some pass overflows even after re-centring
    auto forward = [&](const int *rad, int np, unsigned &mask, double &A) {
        mask = 0;
        A = 1.0; // canonical input
        for (int i = 0; i < np; i++) {
            for (int attempt = 0; attempt < 2; attempt++) {
                double a = attempt ? 0.51 : A;
                bool ok = true;
                for (int s = 0; s < rad[i]; s++) { ok = ok && a < L; a += c(a); }
                ok = ok && a < L; // the canonicalisation / next pass consumes it
                if (ok) { A = a; if (attempt) mask |= 1u << i; break; }
                if (attempt) return false; // even a re-centred pass overflows
                if (i == 0) return false;  // the first pass reads straight from global memory: no re-centring slot
            }
        }
        return true;
    };
    int rad[4];
    const int np = ntt_pass_radices(logN, 0, rad);
    double A;
    if (!forward(rad, np, tb.fwd_recenter, A)) return;
    tb.fwd_recenter_split = 0;
    if (logN == 14) { // CTA-pair form: stage 0 rides on the first pass' loads
        const int split[3] = {6, 4, 4};
        double As;
        if (!forward(split, 3, tb.fwd_recenter_split, As)) return;
        A = std::max(A, As);
    }
    // lazy forward output (|x| <= A p) feeds products of two such values (tensor) or of one with a canonical key word: both operands
    // of a modular product may be lazy only while A*A*p stays below 2^51
    tb.fwd_out_rc = A * A * (double)p >= 0.9 * 2251799813685248.0;
    tb.fwd_out_bound = tb.fwd_out_rc ? 0.51 : A;
    // inverse: canonical or lazy input (tensor output: sum of two fresh products, <= 1.25 p); sums double every stage; bit v
    // re-centres the sums produced by stage v
    A = 1.25;
    for (int v = 0; v < logN; v++) {
        if (2 * A >= L) return;
        const double y = c(2 * A);
        double x = 2 * A;
        if (2 * x >= L) { tb.inv_recenter |= 1u << v; x = 0.51; }
        A = std::max(x, y);
    }
    if (A >= L) return;
    tb.fp_ok = 1;
}
static DigitMap make_digit_map(const std::vector<u64> &q, int w) {
    DigitMap dm;
    memset(&dm, 0, sizeof(dm));
    int d = 0;
    for (int i = 0; i < (int)q.size(); i++) {
        int bits = hm::bit_length(q[i]);
        for (int shift = 0; shift < bits; shift += w) {
            if (d >= 64) throw Error(-1, "decomposition bit count too small: more than 64 digits");
            dm.src[d] = (unsigned char)i;
            dm.shift[d] = (unsigned char)shift;
            d++;
        }
    }
    dm.D = d;
    dm.mask = w >= 64 ? ~0ULL : ((1ULL << w) - 1);
    return dm;
}

Context *context_create(const u64 *plain_primes, int P, uint32_t N, const u64 *coeff, int k, int dbc_relin, int dbc_galois, int device) {
    if (P < 1 || P > 16) throw Error(-1, "need 1..16 plaintext primes");
    int logN = 0;
    while ((1u << logN) < N) logN++;
    if ((1u << logN) != N || logN < 10 || logN > 14) throw Error(-1, "PolyModulusDegree must be a power of two in [1024, 16384]");
    if (k < 1 || k > KMAX) throw Error(-1, "need 1..9 coefficient primes");
    if (dbc_relin < 1 || dbc_relin > 60 || dbc_galois < 1 || dbc_galois > 60) throw Error(-1, "decomposition bit count must be in [1,60]");
    int ndev = 0;
    cudaError_t de = cudaGetDeviceCount(&ndev);
    if (de != cudaSuccess || ndev == 0) throw Error(-2, "libcnhe needs a CUDA device (B200); none is visible -- there is no CPU fallback");
    if (device < 0 || device >= ndev) throw Error(-1, "bad device ordinal");
    std::unique_ptr<Context> cp(new Context());
    Context &c = *cp;
    c.device = device;
    CNHE_CUDA(cudaSetDevice(device));
    c.N = N; c.logN = logN; c.k = k; c.P = P; c.dbc_relin = dbc_relin; c.dbc_galois = dbc_galois;
    c.q.assign(coeff, coeff + k);
    c.t.assign(plain_primes, plain_primes + P);
    for (u64 p : c.q)
        if (!hm::is_prime(p) || (p - 1) % (2ULL * N) || hm::bit_length(p) > 61) throw Error(-1, "coefficient moduli must be primes = 1 mod 2N below 2^61");
    for (u64 p : c.t) {
        if (!hm::is_prime(p) || (p - 1) % (2ULL * N)) throw Error(-1, "plaintext moduli must be primes = 1 mod 2N (batching)");
        for (u64 qq : c.q)
            if (p >= qq) throw Error(-1, "plaintext modulus must be smaller than every coefficient prime");
    }
    // ---- BEHZ base Bsk = auxiliary primes then m_sk.
    // "seal" mode reproduces SEAL 3.2's choice: k auxiliary 61-bit primes (= 1 mod 2^18, descending, after m_sk and gamma)
    // and m_sk = 0x1fffffffffe00001.  The default "fast" mode picks 48-bit primes instead, as many as the exactness bound
    // needs (B*m_sk > 2^8 * N*t*q): every value BEHZ computes is an integer determined by (q, t, m~) alone -- the Bsk residues
    // only carry it -- so the multiply output is bit-identical in both modes (tests/test_gpu_kernels.py checks that), while
    // 48-bit moduli keep every NTT of the multiply on the FP64 butterfly path.
    const u64 M_SK_SEAL = 0x1fffffffffe00001ULL, GAMMA = 0x1fffffffffc80001ULL;
    {
        const char *mode = getenv("CNHE_AUX_BASE");
        const bool seal = mode && std::string(mode) == "seal";
        if (seal) {
            u64 cand = (1ULL << 61) + 1;
            int skipped = 0;
            while ((int)c.bsk.size() < k) {
                cand -= 1ULL << 18;
                if (!hm::is_prime(cand)) continue;
                if (skipped < 2) { skipped++; continue; }
                c.bsk.push_back(cand);
            }
            c.bsk.push_back(M_SK_SEAL);
        } else {
            int need = logN + 8;
            u64 tmax = 0;
            for (u64 t : c.t) tmax = std::max(tmax, t);
            need += hm::bit_length(tmax);
            for (u64 p : c.q) need += hm::bit_length(p);
            const int count = (need + 46) / 47; // each 48-bit prime contributes more than 47 bits
            if (count > KBMAX) throw Error(-1, "parameters too large for the auxiliary base");
            u64 cand = (1ULL << 48) + 1;
            while ((int)c.bsk.size() < std::max(count, 2)) {
                cand -= 2ULL * N;
                if (!hm::is_prime(cand)) continue;
                bool clash = false;
                for (u64 p : c.q) clash = clash || p == cand;
                for (u64 p : c.t) clash = clash || p == cand;
                if (!clash) c.bsk.push_back(cand);
            }
            std::reverse(c.bsk.begin(), c.bsk.end()); // m_sk (last) = the largest
        }
    }
    const int kb = (int)c.bsk.size();
    const u64 M_SK = c.bsk[kb - 1];
    c.kb = kb;
    c.streams.resize(P);
    for (int i = 0; i < P; i++) CNHE_CUDA(cudaStreamCreateWithFlags(&c.streams[i], cudaStreamNonBlocking));
    c.stream = c.streams[0];
    c.trace_ms = trace_slow_ms();
    CNHE_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    CNHE_CUDA(cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
    CNHE_CUDA(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    CNHE_CUDA(cudaEventCreate(&c.ev0));
    CNHE_CUDA(cudaEventCreate(&c.ev1));
    {
        cudaMemPool_t pool;
        CNHE_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = ~0ULL;
        CNHE_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    }
    // ---- NTT tables: ids 0..k-1 q, k..k+kb-1 Bsk, k+kb+c plain modulus c
    const int n_mod = k + kb + P;
    std::vector<u64> moduli;
    for (u64 p : c.q) moduli.push_back(p);
    for (u64 p : c.bsk) moduli.push_back(p);
    for (u64 p : c.t) moduli.push_back(p);
    std::vector<u64> host((size_t)n_mod * 8 * N, 0);
    CNHE_CUDA(cudaMalloc((void **)&c.d_table_mem, host.size() * sizeof(u64)));
    c.h_tabs.resize(n_mod);
    for (int m = 0; m < n_mod; m++) {
        const u64 p = moduli[m];
        const u64 psi = hm::minimal_primitive_root(2ULL * N, p), ipsi = hm::inv(psi, p);
        u64 *w = &host[((size_t)m * 8 + 0) * N], *ws = w + N, *iw = ws + N, *iws = iw + N;
        double *wd = reinterpret_cast<double *>(iws + N), *iwd = wd + N, *wd_hi = iwd + N, *iwd_hi = wd_hi + N;
        u64 a = 1, b = 1;
        for (u64 i = 0; i < N; i++) {
            const u64 r = hm::bit_reverse(i, logN);
            w[r] = a; ws[r] = hm::shoup(a, p);
            iw[r] = b; iws[r] = hm::shoup(b, p);
            wd[r] = a > p / 2 ? -(double)(p - a) : (double)a; // centred, exact below 2^53
            iwd[r] = b > p / 2 ? -(double)(p - b) : (double)b;
            a = hm::mul(a, psi, p);
            b = hm::mul(b, ipsi, p);
        }
        if (logN == 14) { // twiddle tables of the two half-size transforms (see NttTab::fwd_recenter_split)
            const u64 H = N / 2;
            for (u64 h = 0; h < 2; h++)
                for (u64 i = 1; i < H; i++) {
                    u64 m2 = 1;
                    while (2 * m2 <= i) m2 *= 2; // stage of index i: [m2, 2 m2)
                    wd_hi[h * H + i] = wd[i + m2 + h * m2];
                    iwd_hi[h * H + i] = iwd[i + m2 + h * m2];
                }
        } else { // transposed unit-stride twiddles (see NttTab::wd_hi)
            const u64 T = N / 16;
            const int S0 = logN - 4;
            for (u64 j = 0; j < T; j++) {
                for (int u = 0; u < 4; u++)
                    for (int i = 0; i < (1 << u); i++) wd_hi[(u64)((1 << u) - 1 + i) * T + j] = wd[((u64)1 << (S0 + u)) + (j << u) + i];
                for (int i = 0; i < 8; i++) iwd_hi[(u64)i * T + j] = iwd[(N >> 1) + (j << 3) + i];
                for (int i = 0; i < 4; i++) iwd_hi[(u64)(8 + i) * T + j] = iwd[(N >> 2) + (j << 2) + i];
                for (int i = 0; i < 2; i++) iwd_hi[(u64)(12 + i) * T + j] = iwd[(N >> 3) + (j << 1) + i];
                iwd_hi[(u64)14 * T + j] = iwd[(N >> 4) + j];
            }
        }
        NttTab &tb = c.h_tabs[m];
        u64 *base = c.d_table_mem + (size_t)m * 8 * N;
        tb.w = base; tb.ws = base + N; tb.iw = base + 2 * (size_t)N; tb.iws = base + 3 * (size_t)N;
        tb.wd = reinterpret_cast<const double *>(base + 4 * (size_t)N);
        tb.iwd = reinterpret_cast<const double *>(base + 5 * (size_t)N);
        tb.wd_hi = reinterpret_cast<const double *>(base + 6 * (size_t)N);
        tb.iwd_hi = reinterpret_cast<const double *>(base + 7 * (size_t)N);
        tb.inv_n = hm::inv(N % p, p);
        tb.inv_n_s = hm::shoup(tb.inv_n, p);
        tb.mod = make_dmod(p);
        tb.pd = (double)p;
        tb.pinv = 1.0 / (double)p;
        tb.inv_n_d = tb.inv_n > p / 2 ? -(double)(p - tb.inv_n) : (double)tb.inv_n;
        {
            const u64 nw = hm::mul(iw[1], tb.inv_n, p);
            tb.inv_n_w_d = nw > p / 2 ? -(double)(p - nw) : (double)nw;
        }
        fp_schedule(tb, logN, getenv("CNHE_NTT_INT") != nullptr);
    }
    CNHE_CUDA(cudaMemcpy(c.d_table_mem, host.data(), host.size() * sizeof(u64), cudaMemcpyHostToDevice));
    CNHE_CUDA(cudaMalloc((void **)&c.d_tabs, n_mod * sizeof(NttTab)));
    CNHE_CUDA(cudaMemcpy(c.d_tabs, c.h_tabs.data(), n_mod * sizeof(NttTab), cudaMemcpyHostToDevice));
    // ---- BatchEncoder index map (BatchEncoder::populate_matrix_reps_index_map)
    c.h_index_map.resize(N);
    {
        const u64 row = N >> 1, m2 = 2ULL * N;
        u64 pos = 1;
        for (u64 i = 0; i < row; i++) {
            c.h_index_map[i] = (u32)hm::bit_reverse((pos - 1) >> 1, logN);
            c.h_index_map[row | i] = (u32)hm::bit_reverse((m2 - pos - 1) >> 1, logN);
            pos = (pos * 3) & (m2 - 1);
        }
    }
    CNHE_CUDA(cudaMalloc((void **)&c.d_index_map, N * sizeof(u32)));
    CNHE_CUDA(cudaMemcpy(c.d_index_map, c.h_index_map.data(), N * sizeof(u32), cudaMemcpyHostToDevice));
    // ---- BEHZ constants (BaseConverter::generate)
    BehzConst &bc = c.h_bc;
    memset(&bc, 0, sizeof(bc));
    bc.k = k;
    bc.kb = kb;
    bc.centered_mtilde = 0;
    const int na = kb - 1;
    std::vector<u64> B(c.bsk.begin(), c.bsk.begin() + na);
    for (int i = 0; i < k; i++) bc.q[i] = make_dmod(c.q[i]);
    for (int j = 0; j < kb; j++) bc.bsk[j] = make_dmod(c.bsk[j]);
    const u64 MT = 1ULL << 32;
    u64 q_mod_mt = 1;
    for (int i = 0; i < k; i++) q_mod_mt = (q_mod_mt * (c.q[i] & 0xffffffffULL)) & 0xffffffffULL;
    {
        u64 x = q_mod_mt; // Newton iteration for the inverse modulo 2^32 (q is odd)
        for (int it = 0; it < 6; it++) x *= 2 - q_mod_mt * x;
        bc.inv_q_mod_mtilde = x & 0xffffffffULL;
    }
    for (int i = 0; i < k; i++) {
        const u64 qi = c.q[i];
        const u64 inv = hm::inv(hm::product_mod(c.q, i, qi), qi);
        bc.inv_qhat_mod_q[i] = inv;
        bc.mtilde_inv_qhat_mod_q[i] = hm::mul(inv, MT % qi, qi);
        u64 pm = 1;
        for (int l = 0; l < k; l++)
            if (l != i) pm = (pm * (c.q[l] & 0xffffffffULL)) & 0xffffffffULL;
        bc.qhat_mod_mtilde[i] = pm;
        bc.B_mod_q[i] = hm::product_mod(B, -1, qi);
        for (int j = 0; j < na; j++) bc.bhat_mod_q[i][j] = hm::product_mod(B, j, qi);
    }
    for (int j = 0; j < kb; j++) {
        const u64 bj = c.bsk[j];
        for (int i = 0; i < k; i++) bc.qhat_mod_bsk[j][i] = hm::product_mod(c.q, i, bj);
        bc.q_mod_bsk[j] = hm::product_mod(c.q, -1, bj);
        bc.inv_q_mod_bsk[j] = hm::inv(bc.q_mod_bsk[j], bj);
        bc.inv_mtilde_mod_bsk[j] = hm::inv(MT % bj, bj);
    }
    for (int j = 0; j < na; j++) {
        bc.inv_bhat_mod_b[j] = hm::inv(hm::product_mod(B, j, B[j]), B[j]);
        bc.bhat_mod_msk[j] = hm::product_mod(B, j, M_SK);
    }
    bc.inv_B_mod_msk = hm::inv(hm::product_mod(B, -1, M_SK), M_SK);
    CNHE_CUDA(cudaMalloc((void **)&c.d_bc, sizeof(BehzConst)));
    CNHE_CUDA(cudaMemcpy(c.d_bc, &bc, sizeof(BehzConst), cudaMemcpyHostToDevice));
    // FP64 twin of the constants (used when every q_i and Bsk prime is below 2^50)
    {
        BehzConstF &f = c.h_bf;
        memset(&f, 0, sizeof(f));
        auto cen = [](u64 v, u64 p) { return v > p / 2 ? -(double)(p - v) : (double)v; };
        f.k = k; f.kb = kb; f.centered_mtilde = 0;
        c.fp_elementwise = getenv("CNHE_NTT_INT") == nullptr;
        for (int i = 0; i < k; i++) {
            const u64 p = c.q[i];
            c.fp_elementwise = c.fp_elementwise && hm::bit_length(p) <= 49;
            f.qd[i] = (double)p; f.qinv[i] = 1.0 / (double)p; f.q_u[i] = p;
            f.inv_qhat_mod_q[i] = cen(bc.inv_qhat_mod_q[i], p);
            f.mtilde_inv_qhat_mod_q[i] = cen(bc.mtilde_inv_qhat_mod_q[i], p);
            f.qhat_mod_mtilde[i] = bc.qhat_mod_mtilde[i];
            f.B_mod_q[i] = cen(bc.B_mod_q[i], p);
            for (int j = 0; j < na; j++) f.bhat_mod_q[i][j] = cen(bc.bhat_mod_q[i][j], p);
        }
        f.inv_q_mod_mtilde = bc.inv_q_mod_mtilde;
        for (int j = 0; j < kb; j++) {
            const u64 p = c.bsk[j];
            c.fp_elementwise = c.fp_elementwise && hm::bit_length(p) <= 48;
            f.bd[j] = (double)p; f.binv[j] = 1.0 / (double)p; f.b_u[j] = p;
            for (int i = 0; i < k; i++) f.qhat_mod_bsk[j][i] = cen(bc.qhat_mod_bsk[j][i], p);
            f.q_mod_bsk[j] = cen(bc.q_mod_bsk[j], p);
            f.inv_q_mod_bsk[j] = cen(bc.inv_q_mod_bsk[j], p);
            f.inv_mtilde_mod_bsk[j] = cen(bc.inv_mtilde_mod_bsk[j], p);
        }
        for (int j = 0; j < na; j++) {
            f.inv_bhat_mod_b[j] = cen(bc.inv_bhat_mod_b[j], c.bsk[j]);
            f.bhat_mod_msk[j] = cen(bc.bhat_mod_msk[j], M_SK);
        }
        f.inv_B_mod_msk = cen(bc.inv_B_mod_msk, M_SK);
        f.msk_half = (double)(M_SK >> 1);
        // lazy-double intermediates between the kernels of a multiply / key switch: FP64 everywhere on the q and Bsk transforms
        c.lazy = c.fp_elementwise && getenv("CNHE_NO_LAZY") == nullptr;
        for (int i = 0; i < k + kb; i++) c.lazy = c.lazy && c.h_tabs[i].fp_ok;
        CNHE_CUDA(cudaMalloc((void **)&c.d_bf, sizeof(BehzConstF)));
        CNHE_CUDA(cudaMemcpy(c.d_bf, &f, sizeof(BehzConstF), cudaMemcpyHostToDevice));
    }
    // ---- per plaintext modulus
    c.ch.resize(P);
    for (int ci = 0; ci < P; ci++) {
        Channel &ch = c.ch[ci];
        const u64 t = c.t[ci];
        ch.t = t;
        rng_from_os(ch.rng); // encryption randomness of a context that only imports a public key is unpredictable too
        ch.mod_id = k + kb + ci;
        PlainConst &pc = ch.pc;
        memset(&pc, 0, sizeof(pc));
        pc.t = t;
        pc.threshold = (t + 1) >> 1;
        pc.gamma = GAMMA;
        pc.tmod = make_dmod(t);
        pc.gmod = make_dmod(GAMMA);
        std::vector<u64> quot;
        u64 rem;
        hm::div_product(c.q, t, quot, rem);
        for (int i = 0; i < k; i++) {
            pc.delta[i] = hm::limbs_mod(quot, c.q[i]);
            pc.q_mod_t[i] = rem % c.q[i];
            pc.tgamma_mod_q[i] = hm::mul(t % c.q[i], GAMMA % c.q[i], c.q[i]);
            pc.qhat_mod_t[i] = hm::product_mod(c.q, i, t);
            pc.qhat_mod_gamma[i] = hm::product_mod(c.q, i, GAMMA);
        }
        { // folded constants of the floor kernel (FloorConstF)
            FloorConstF &ff = ch.floor_f;
            memset(&ff, 0, sizeof(ff));
            auto cen = [](u64 v, u64 p) { return v > p / 2 ? -(double)(p - v) : (double)v; };
            ff.k = k; ff.kb = kb;
            const int na_ = kb - 1;
            for (int i = 0; i < k; i++) {
                const u64 p = c.q[i];
                ff.qd[i] = (double)p; ff.qinv[i] = 1.0 / (double)p;
                ff.xq[i] = cen(hm::mul(t % p, bc.inv_qhat_mod_q[i], p), p);
                ff.B_mod_q[i] = cen(bc.B_mod_q[i], p);
                for (int j = 0; j < na_; j++) ff.bhat_mod_q[i][j] = cen(bc.bhat_mod_q[i][j], p);
            }
            for (int j = 0; j < kb; j++) {
                const u64 p = c.bsk[j];
                ff.bd[j] = (double)p; ff.binv[j] = 1.0 / (double)p;
                u64 post = bc.inv_q_mod_bsk[j];                                   // q^-1 mod p_j ...
                if (j < na_) post = hm::mul(post, bc.inv_bhat_mod_b[j], p);       // ... times B-hat_j^-1 for the base-B primes
                ff.xb[j] = cen(hm::mul(t % p, post, p), p);
                for (int i = 0; i < k; i++) ff.conv[j][i] = cen(hm::mul(bc.qhat_mod_bsk[j][i], post, p), p);
            }
            for (int j = 0; j < na_; j++) ff.bhat_mod_msk[j] = cen(bc.bhat_mod_msk[j], M_SK);
            ff.inv_B_mod_msk = cen(bc.inv_B_mod_msk, M_SK);
            ff.msk_half = (double)(M_SK >> 1);
        }
        pc.neg_inv_q_mod_t = hm::neg(hm::inv(hm::product_mod(c.q, -1, t), t), t);
        pc.neg_inv_q_mod_gamma = hm::neg(hm::inv(hm::product_mod(c.q, -1, GAMMA), GAMMA), GAMMA);
        pc.inv_gamma_mod_t = hm::inv(GAMMA % t, t);
    }
    c.dm_relin = make_digit_map(c.q, dbc_relin);
    c.dm_galois = make_digit_map(c.q, dbc_galois);
    // Galois elements of KeyGenerator::galois_keys(dbc): 2N-1, then 3^(2^i), 3^-(2^i) for i < logN-1
    {
        const u64 m2 = 2ULL * N;
        c.galois_elts.push_back(m2 - 1);
        u64 p3 = 3, n3 = 0;
        for (u64 x = 1; x < m2; x += 2)
            if (((x * 3) & (m2 - 1)) == 1) { n3 = x; break; }
        for (int i = 0; i < logN - 1; i++) {
            c.galois_elts.push_back(p3);
            p3 = (p3 * p3) & (m2 - 1);
            c.galois_elts.push_back(n3);
            n3 = (n3 * n3) & (m2 - 1);
        }
    }
    // ---- wrapper CRT data (EncryptedSealBfvEnvironment.PreCompute, "EncryptedSealBfvVector.cs:79-90")
    {
        unsigned __int128 big = 1;
        int bits = 0;
        for (u64 t : c.t) bits += hm::bit_length(t);
        if (bits > 126) throw Error(-1, "product of the plaintext moduli must stay below 2^126");
        for (u64 t : c.t) big *= t;
        c.big_factor = big;
        for (int i = 0; i < P; i++) {
            unsigned __int128 minor = big / c.t[i];
            u64 y = hm::inv((u64)(minor % c.t[i]), c.t[i]);
            c.crt_coeff.push_back(minor * y);
        }
    }
    CNHE_CUDA(cudaDeviceSynchronize());
    return cp.release();
}

// ---------------------------------------------------------------- pointer tables
const u64 *const *upload_ptrs(Context &c, const std::vector<const u64 *> &ptrs) {
    u64 *d = c.ws_alloc(ptrs.size());
    c.h2d(d, ptrs.data(), ptrs.size() * sizeof(u64 *));
    return reinterpret_cast<const u64 *const *>(d);
}
u64 *const *upload_ptrs_mut(Context &c, const std::vector<u64 *> &ptrs) {
    u64 *d = c.ws_alloc(ptrs.size());
    c.h2d(d, ptrs.data(), ptrs.size() * sizeof(u64 *));
    return reinterpret_cast<u64 *const *>(d);
}

// ---------------------------------------------------------------- core operations
static int fp_range(const Context &c, int mod_base, int mod_count) {
    for (int i = mod_base; i < mod_base + mod_count; i++)
        if (!c.h_tabs[i].fp_ok) return 0;
    return 1;
}
void op_ntt(Context &c, const u64 *src, u64 *dst, int n_polys, int mod_base, int mod_count, bool inverse) {
    PROF(inverse ? 1 : 0, 16.0 * c.N * n_polys);
    c.check(inverse ? launch_ntt_inverse(src, dst, n_polys, c.logN, c.d_tabs, mod_base, mod_count, fp_range(c, mod_base, mod_count), c.stream)
                    : launch_ntt_forward(src, dst, n_polys, c.logN, c.d_tabs, mod_base, mod_count, fp_range(c, mod_base, mod_count), c.stream),
            "ntt");
}

// out[i] = (base_i + sum_d NTT^-1(NTT(digit_d(target_i)) * key_d)): ciphertext i's target polynomial (k residues) is at
// target + i * target_stride, its base (2 polynomials) at base + i * base_stride; out is packed [n][2][k][N]
void op_key_switch(Context &c, const u64 *target, size_t target_stride, int n, const u64 *key, const DigitMap &dm, const u64 *base,
                   size_t base_stride, u64 *out) {
    const int k = c.k;
    const size_t N = c.N;
    const int fpq = fp_range(c, 0, k);
    const bool lazy = c.lazy && fpq;
    const int wave = c.wave(((size_t)dm.D * k + 2 * k) * N);
    for (int c0 = 0; c0 < n; c0 += wave) {
        WsScope scope(c);
        const int m = std::min(wave, n - c0);
        u64 *digits = c.ws_alloc((size_t)m * dm.D * k * N);
        u64 *acc = c.ws_alloc((size_t)m * 2 * k * N);
        {
            PROF(0, 16.0 * N * (double)m * dm.D * k); // 16N bytes per transform (8N digit source read + 8N written)
            c.check(launch_ntt_forward_digits(target + (size_t)c0 * target_stride, target_stride, digits, m, k, dm, c.logN, c.d_tabs,
                                              fpq | (lazy ? NTT_OUT_F : 0), c.stream),
                    "ntt_forward_digits");
        }
        {
            PROF(3, 8.0 * N * ((double)m * dm.D * k + (double)dm.D * 2 * k + (double)m * 2 * k));
            if (c.fp_elementwise) c.check(launch_ks_mac_fp(digits, key, acc, m, dm.D, k, c.logN, &c.h_bf, lazy, c.stream), "ks_mac_fp");
            else c.check(launch_ks_mac(digits, key, acc, m, dm.D, k, c.logN, c.d_bc, c.stream), "ks_mac");
        }
        PROF(1, 24.0 * N * m * 2 * k);
        c.check(launch_ntt_inverse_add(acc, base + (size_t)c0 * base_stride, 2 * k, base_stride, out + (size_t)c0 * 2 * k * N, m * 2 * k, c.logN,
                                       c.d_tabs, 0, k, fpq | (lazy ? NTT_IN_F : 0), c.stream),
                "ntt_inverse_add");
    }
}

static void multiply_chunk(Context &c, int ch, const std::vector<const u64 *> &a, const std::vector<const u64 *> &b, int c0, int m, u64 *out3) {
    const int k = c.k, kt = k + c.kb;
    const size_t N = c.N;
    bool square = true;
    for (int i = 0; i < m; i++) square = square && a[c0 + i] == b[c0 + i];
    std::vector<const u64 *> pa(a.begin() + c0, a.begin() + c0 + m);
    u64 *A = c.ws_alloc((size_t)m * 2 * kt * N);
    const int fpt = fp_range(c, 0, kt);
    const int lazy = c.lazy && fpt ? 1 : 0;
    const int fmt = fpt | (lazy ? NTT_IN_F | NTT_OUT_F : 0);
    {
        PROF(2, 8.0 * N * m * 2 * (k + kt));
        if (c.fp_elementwise) c.check(launch_behz_lift_fp(upload_ptrs(c, pa), A, m, c.logN, &c.h_bf, lazy, c.stream), "behz_lift_fp");
        else c.check(launch_behz_lift(upload_ptrs(c, pa), A, m, c.logN, c.d_bc, c.stream), "behz_lift");
    }
    {
        PROF(0, 16.0 * N * m * 2 * kt);
        c.check(launch_ntt_forward(A, A, m * 2 * kt, c.logN, c.d_tabs, 0, kt, fmt, c.stream), "ntt_forward");
    }
    u64 *B = A;
    if (!square) {
        std::vector<const u64 *> pb(b.begin() + c0, b.begin() + c0 + m);
        B = c.ws_alloc((size_t)m * 2 * kt * N);
        if (c.fp_elementwise) c.check(launch_behz_lift_fp(upload_ptrs(c, pb), B, m, c.logN, &c.h_bf, lazy, c.stream), "behz_lift_fp");
        else c.check(launch_behz_lift(upload_ptrs(c, pb), B, m, c.logN, c.d_bc, c.stream), "behz_lift");
        c.check(launch_ntt_forward(B, B, m * 2 * kt, c.logN, c.d_tabs, 0, kt, fmt, c.stream), "ntt_forward");
    }
    u64 *D = c.ws_alloc((size_t)m * 3 * kt * N);
    {
        PROF(2, 8.0 * N * m * kt * (square ? 5 : 7));
        if (c.fp_elementwise) c.check(launch_behz_tensor_fp(A, B, D, m, kt, c.logN, &c.h_bf, lazy, c.stream), "behz_tensor_fp");
        else c.check(launch_behz_tensor(A, B, D, m, kt, c.logN, c.d_bc, c.stream), "behz_tensor");
    }
    {
        PROF(1, 16.0 * N * m * 3 * kt);
        c.check(launch_ntt_inverse(D, D, m * 3 * kt, c.logN, c.d_tabs, 0, kt, fmt, c.stream), "ntt_inverse");
    }
    PROF(2, 8.0 * N * m * 3 * (kt + k));
    static const bool fold = getenv("CNHE_FLOOR_NOFOLD") == nullptr;
    if (c.fp_elementwise && lazy && fold) c.check(launch_behz_floor_fold_fp(D, out3, m, c.logN, &c.ch[ch].floor_f, c.stream), "behz_floor_fold_fp");
    else if (c.fp_elementwise) c.check(launch_behz_floor_fp(D, out3, m, c.ch[ch].t, c.logN, &c.h_bf, lazy, c.stream), "behz_floor_fp");
    else c.check(launch_behz_floor(D, out3, m, c.ch[ch].t, c.logN, c.d_bc, c.stream), "behz_floor");
}
void op_multiply(Context &c, int ch, const std::vector<const u64 *> &a, const std::vector<const u64 *> &b, u64 *out3) {
    const int n = (int)a.size();
    const int wave = c.wave((size_t)(7 * (c.k + c.kb)) * c.N);
    for (int c0 = 0; c0 < n; c0 += wave) {
        WsScope scope(c);
        const int m = std::min(wave, n - c0);
        multiply_chunk(c, ch, a, b, c0, m, out3 + (size_t)c0 * 3 * c.k * c.N);
    }
    c.note(Context::OP_MULTIPLY, ch, n);
}
void op_relinearize(Context &c, int ch, const u64 *in3, int n, u64 *out2) {
    if (!c.ch[ch].have_rlk) throw Error(-3, "relinearization keys are missing");
    const int k = c.k;
    const size_t N = c.N;
    // the size-3 layout [c0 c1 c2] is consumed in place: c2 is the key-switch target, (c0, c1) the base it is added to
    const size_t s3 = (size_t)3 * k * N;
    op_key_switch(c, in3 + (size_t)2 * k * N, s3, n, c.ch[ch].rlk->p, c.dm_relin, in3, s3, out2);
    c.note(Context::OP_RELINEARIZE, ch, n, out2);
}
void op_multiply_relin(Context &c, int ch, const std::vector<const u64 *> &a, const std::vector<const u64 *> &b, u64 *out2) {
    if (!c.ch[ch].have_rlk) throw Error(-3, "relinearization keys are missing");
    const int n = (int)a.size(), k = c.k;
    const size_t N = c.N;
    const int wave = c.wave(((size_t)c.dm_relin.D * k + 7 * (k + c.kb) + 5 * k) * N);
    for (int c0 = 0; c0 < n; c0 += wave) {
        WsScope scope(c); // stream-ordered frees: the next wave reuses the memory once these kernels are done
        const int m = std::min(wave, n - c0);
        u64 *ct3 = c.ws_alloc((size_t)m * 3 * k * N);
        multiply_chunk(c, ch, a, b, c0, m, ct3);
        const size_t s3 = (size_t)3 * k * N;
        op_key_switch(c, ct3 + (size_t)2 * k * N, s3, m, c.ch[ch].rlk->p, c.dm_relin, ct3, s3, out2 + (size_t)c0 * 2 * k * N);
    }
    c.op_count[Context::OP_MULTIPLY] += (uint64_t)n;
    c.note(Context::OP_RELINEARIZE, ch, n, out2, a[0], b[0]);
}

u64 galois_elt_from_step(const Context &c, int steps) { // Evaluator::galois_elt_from_step: positive = rotate left
    const u64 n = c.N, m = 2 * n;
    if (steps == 0) return m - 1;
    const bool neg = steps < 0;
    const u64 pos = neg ? (u64)(-(long long)steps) : (u64)steps;
    if (pos >= (n >> 1)) throw Error(-1, "step count too large");
    const u64 s = neg ? (n >> 1) - pos : pos;
    u64 e = 1;
    for (u64 i = 0; i < s; i++) e = (e * 3) & (m - 1);
    return e;
}
void op_apply_galois(Context &c, int ch, const u64 *in, int n, u64 elt, u64 *out, bool add_back) {
    auto it = c.ch[ch].glk.find(elt);
    if (it == c.ch[ch].glk.end()) throw Error(-3, "Galois key not present");
    const int k = c.k;
    const size_t N = c.N;
    const u64 m2 = 2ULL * N;
    u64 einv = 0;
    for (u64 x = 1; x < m2; x += 2)
        if (((x * elt) & (m2 - 1)) == 1) { einv = x; break; }
    for (int c0 = 0; c0 < n; c0 += c.chunk) {
        const int m = std::min(c.chunk, n - c0);
        u64 *base = c.ws_alloc((size_t)m * 2 * k * N), *p1 = c.ws_alloc((size_t)m * k * N);
        c.check(launch_galois(in + (size_t)c0 * 2 * k * N, base, p1, m, einv, k, c.logN, c.d_bc, c.stream, add_back ? 1 : 0), "galois");
        op_key_switch(c, p1, (size_t)k * N, m, it->second->p, c.dm_galois, base, (size_t)2 * k * N, out + (size_t)c0 * 2 * k * N);
    }
    c.note(elt == m2 - 1 ? Context::OP_ROTATE_COLUMNS : Context::OP_ROTATE_ROWS_HOP, ch, n, out, in);
    if (add_back) c.note(Context::OP_ADD, ch, n, out, in, out); // the reference issues Rotate + Add: both are counted
}
// x + rotate(x) in one pass (in == out allowed): the permutation kernel folds the unrotated ciphertext into the key switch's base.  Returns
// false when the step has no key of its own (multi-hop rotation) or per-operation noise tracing wants the rotated ciphertext on its own.
bool op_rotate_add(Context &c, int ch, const u64 *in, int n, int steps, bool columns, u64 *out) {
    if (c.trace_noise) return false;
    const u64 elt = columns ? 2ULL * c.N - 1 : galois_elt_from_step(c, steps);
    if (!c.ch[ch].glk.count(elt)) return false;
    op_apply_galois(c, ch, in, n, elt, out, true);
    return true;
}
static std::vector<int> naf(int value) { // non-adjacent form, least significant term first (SEAL util::naf)
    std::vector<int> res;
    const bool sign = value < 0;
    int v = sign ? -value : value;
    for (int i = 0; v; i++) {
        const int zi = (v & 1) ? 2 - (v & 3) : 0;
        v = (v - zi) >> 1;
        if (zi) res.push_back((sign ? -zi : zi) * (1 << i));
    }
    return res;
}
void op_rotate_rows(Context &c, int ch, const u64 *in, int n, int steps, u64 *out) { // Evaluator::rotate_internal
    const size_t words = (size_t)n * c.ct_words();
    if (steps == 0) {
        if (in != out) { CNHE_CUDA(cudaMemcpyAsync(out, in, words * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(out, in); }
        return;
    }
    const u64 elt = galois_elt_from_step(c, steps);
    if (c.ch[ch].glk.count(elt)) { op_apply_galois(c, ch, in, n, elt, out); return; }
    std::vector<int> hops = naf(steps);
    if (hops.size() == 1) throw Error(-3, "Galois key not present");
    const u64 *cur = in;
    for (size_t h = 0; h < hops.size(); h++) {
        if ((size_t)std::abs(hops[h]) == (c.N >> 1)) continue;
        u64 *nxt = (h + 1 == hops.size()) ? out : c.ws_alloc(words);
        op_rotate_rows(c, ch, cur, n, hops[h], nxt);
        cur = nxt;
    }
    if (cur != out) { CNHE_CUDA(cudaMemcpyAsync(out, cur, words * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(out, cur); }
}
void op_rotate_columns(Context &c, int ch, const u64 *in, int n, u64 *out) { op_apply_galois(c, ch, in, n, 2ULL * c.N - 1, out); }

// apply_galois on n ciphertexts scattered in memory (pointer table) -> packed out[n]
static void apply_galois_gather(Context &c, int ch, const std::vector<const u64 *> &ins, u64 elt, u64 *out) {
    auto it = c.ch[ch].glk.find(elt);
    if (it == c.ch[ch].glk.end()) throw Error(-3, "Galois key not present");
    const int k = c.k, n = (int)ins.size();
    const size_t N = c.N;
    const u64 m2 = 2ULL * N;
    u64 einv = 0;
    for (u64 x = 1; x < m2; x += 2)
        if (((x * elt) & (m2 - 1)) == 1) { einv = x; break; }
    for (int c0 = 0; c0 < n; c0 += c.chunk) {
        const int m = std::min(c.chunk, n - c0);
        std::vector<const u64 *> part(ins.begin() + c0, ins.begin() + c0 + m);
        u64 *base = c.ws_alloc((size_t)m * 2 * k * N), *p1 = c.ws_alloc((size_t)m * k * N);
        c.check(launch_galois_gather(upload_ptrs(c, part), base, p1, m, einv, k, c.logN, c.d_bc, c.stream), "galois");
        op_key_switch(c, p1, (size_t)k * N, m, it->second->p, c.dm_galois, base, (size_t)2 * k * N, out + (size_t)c0 * 2 * k * N);
    }
    c.op_count[elt == m2 - 1 ? Context::OP_ROTATE_COLUMNS : Context::OP_ROTATE_ROWS_HOP] += (uint64_t)n;
    if (c.trace_noise) // one record per ciphertext, as the unbatched path would have written
        for (int i = 0; i < n; i++) {
            c.op_count[Context::OP_ROTATE_ROWS_HOP] -= 1; // note() counts it again
            c.note(Context::OP_ROTATE_ROWS_HOP, ch, 1, out + (size_t)i * 2 * k * N, ins[i]);
        }
}
void op_rotate_rows_multi(Context &c, int ch, const std::vector<RotateJob> &jobs) {
    const size_t ctw = c.ct_words();
    struct State { std::vector<int> hops; size_t next; const u64 *cur; };
    std::vector<State> st(jobs.size());
    for (size_t j = 0; j < jobs.size(); j++) {
        const RotateJob &job = jobs[j];
        st[j].next = 0;
        st[j].cur = job.src;
        if (job.steps == 0) continue;
        const u64 elt = galois_elt_from_step(c, job.steps);
        if (c.ch[ch].glk.count(elt)) st[j].hops = {job.steps};
        else {
            for (int h : naf(job.steps))
                if ((size_t)std::abs(h) != (c.N >> 1)) st[j].hops.push_back(h); // rotate_internal skips a hop of exactly N/2
            if (naf(job.steps).size() == 1) throw Error(-3, "Galois key not present");
        }
    }
    for (;;) {
        // the hop value most jobs are waiting for next
        std::map<int, std::vector<size_t>> want;
        for (size_t j = 0; j < jobs.size(); j++)
            if (st[j].next < st[j].hops.size()) want[st[j].hops[st[j].next]].push_back(j);
        if (want.empty()) break;
        auto best = want.begin();
        for (auto it = want.begin(); it != want.end(); ++it)
            if (it->second.size() > best->second.size()) best = it;
        const std::vector<size_t> &js = best->second;
        std::vector<const u64 *> ins;
        for (size_t j : js) ins.push_back(st[j].cur);
        u64 *out = c.ws_alloc(js.size() * ctw);
        apply_galois_gather(c, ch, ins, galois_elt_from_step(c, best->first), out);
        for (size_t i = 0; i < js.size(); i++) {
            st[js[i]].cur = out + i * ctw;
            st[js[i]].next++;
        }
    }
    for (size_t j = 0; j < jobs.size(); j++)
        if (st[j].cur != jobs[j].dst) {
            CNHE_CUDA(cudaMemcpyAsync(jobs[j].dst, st[j].cur, ctw * 8, cudaMemcpyDeviceToDevice, c.stream));
            c.note_copy(jobs[j].dst, st[j].cur);
        }
}

void op_multiply_plain_dense(Context &c, int ch, const u64 *ct, int n, const u64 *plain, bool plain_per_ct, u64 *out) {
    const int k = c.k;
    const size_t N = c.N;
    const int np = plain_per_ct ? n : 1;
    u64 *lifted = c.ws_alloc((size_t)np * k * N);
    c.check(launch_plain_lift(plain, lifted, np, (int)N, k, c.logN, c.d_bc, c.ch[ch].pc, c.stream), "plain_lift");
    c.check(launch_ntt_forward(lifted, lifted, np * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
    u64 *tmp = c.ws_alloc((size_t)n * 2 * k * N);
    c.check(launch_ntt_forward(ct, tmp, n * 2 * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
    c.check(launch_dyadic_bcast(tmp, lifted, tmp, n, 2, 1, plain_per_ct ? 1 : 0, k, c.logN, c.d_bc, c.stream), "dyadic");
    c.check(launch_ntt_inverse(tmp, out, n * 2 * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_inverse");
    c.note(Context::OP_MULTIPLY_PLAIN, ch, n, out, ct);
}
// one ciphertext times n dense plaintexts: out[i] = ct * plain[i]   (row-major matrix x vector: every row against the same input)
void op_multiply_plain_dense_bcast(Context &c, int ch, const u64 *ct, const u64 *plains, int n, u64 *out) {
    const int k = c.k;
    const size_t N = c.N;
    u64 *ctn = c.ws_alloc((size_t)2 * k * N);
    c.check(launch_ntt_forward(ct, ctn, 2 * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
    for (int c0 = 0; c0 < n; c0 += 4 * c.chunk) {
        WsScope scope(c);
        const int m = std::min(4 * c.chunk, n - c0);
        u64 *lifted = c.ws_alloc((size_t)m * k * N);
        c.check(launch_plain_lift(plains + (size_t)c0 * N, lifted, m, (int)N, k, c.logN, c.d_bc, c.ch[ch].pc, c.stream), "plain_lift");
        c.check(launch_ntt_forward(lifted, lifted, m * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
        u64 *dst = out + (size_t)c0 * 2 * k * N;
        c.check(launch_dyadic_bcast(ctn, lifted, dst, m, 2, 0, 1, k, c.logN, c.d_bc, c.stream), "dyadic");
        c.check(launch_ntt_inverse(dst, dst, m * 2 * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_inverse");
    }
    c.note(Context::OP_MULTIPLY_PLAIN, ch, n, out, ct);
}
void op_encode(Context &c, int ch, const u64 *values, int n, int count, u64 *plain) {
    c.check(launch_encode_scatter(values, plain, n, count, c.d_index_map, c.logN, c.stream), "encode_scatter");
    c.check(launch_ntt_inverse(plain, plain, n, c.logN, c.d_tabs, c.ch[ch].mod_id, 1, fp_range(c, c.ch[ch].mod_id, 1), c.stream), "ntt_inverse(t)");
}
void op_encode_onehot(Context &c, int ch, int n, int first_col, u64 *plain) {
    c.check(launch_onehot_scatter(plain, n, first_col, c.d_index_map, c.logN, c.stream), "onehot_scatter");
    c.check(launch_ntt_inverse(plain, plain, n, c.logN, c.d_tabs, c.ch[ch].mod_id, 1, fp_range(c, c.ch[ch].mod_id, 1), c.stream), "ntt_inverse(t)");
}
void op_decode(Context &c, int ch, const u64 *plain, int n, u64 *values) {
    u64 *tmp = c.ws_alloc((size_t)n * c.N);
    c.check(launch_ntt_forward(plain, tmp, n, c.logN, c.d_tabs, c.ch[ch].mod_id, 1, fp_range(c, c.ch[ch].mod_id, 1), c.stream), "ntt_forward(t)");
    c.check(launch_decode_gather(tmp, values, n, c.d_index_map, c.logN, c.stream), "decode_gather");
}
u64 take_nonces(Context &c, int chi, u64 n) {
    Channel &ch = c.ch[chi];
    if (n >= (1ULL << 31)) throw Error(-1, "too many encryptions in one call");
    if (ch.nonce + n >= (1ULL << 32)) { // the stream id carries 32 bits of the counter
        if (!ch.rng.secure) throw Error(-1, "deterministic (seeded, test-only) channel exhausted its 2^32 encryption nonces");
        rng_from_os(ch.rng);
        ch.nonce = 1;
    }
    const u64 n0 = ch.nonce;
    ch.nonce += n;
    return n0;
}
void op_encrypt(Context &c, int chi, const u64 *plain, size_t plain_stride, int n, int coeffs, u64 nonce0, u64 *ct) {
    Channel &ch = c.ch[chi];
    if (!ch.have_pk) throw Error(-3, "public key is missing");
    const int k = c.k;
    const size_t N = c.N;
    for (int c0 = 0; c0 < n; c0 += 4 * c.chunk) {
        const int m = std::min(4 * c.chunk, n - c0);
        u64 *u = c.ws_alloc((size_t)m * k * N);
        c.check(launch_sample(u, m, SAMPLE_TERNARY, ch.rng, stream_id(8, nonce0 + c0, 0), 1ULL << 16, k, c.logN, c.d_bc, c.stream), "sample");
        c.check(launch_ntt_forward(u, u, m * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
        u64 *dst = ct + (size_t)c0 * 2 * k * N;
        c.check(launch_dyadic_bcast(ch.pk->p, u, dst, m, 2, 0, 1, k, c.logN, c.d_bc, c.stream), "dyadic");
        c.check(launch_ntt_inverse(dst, dst, m * 2 * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_inverse");
        c.check(launch_encrypt_finish(dst, plain ? plain + (size_t)c0 * plain_stride : nullptr, plain_stride, m, plain ? coeffs : 0, ch.rng,
                                      nonce0 + c0, k, c.logN, c.d_bc, ch.pc, c.stream),
                "encrypt_finish");
    }
    c.note(Context::OP_ENCRYPT, chi, n, ct);
}
static void dot_with_secret(Context &c, int chi, const u64 *ct, int n, u64 *x) {
    Channel &ch = c.ch[chi];
    if (!ch.have_sk) throw Error(-3, "secret key is missing");
    const int k = c.k;
    const size_t N = c.N, kN = (size_t)k * N;
    u64 *c0 = c.ws_alloc((size_t)n * kN), *c1 = c.ws_alloc((size_t)n * kN);
    CNHE_CUDA(cudaMemcpy2DAsync(c0, kN * 8, ct, 2 * kN * 8, kN * 8, n, cudaMemcpyDeviceToDevice, c.stream));
    CNHE_CUDA(cudaMemcpy2DAsync(c1, kN * 8, ct + kN, 2 * kN * 8, kN * 8, n, cudaMemcpyDeviceToDevice, c.stream));
    c.check(launch_ntt_forward(c1, c1, n * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
    c.check(launch_dyadic_bcast(c1, ch.sk->p, c1, n, 1, 1, 0, k, c.logN, c.d_bc, c.stream), "dyadic");
    c.check(launch_ntt_inverse_add(c1, c0, 1, N, x, n * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_inverse_add");
}
void op_decrypt(Context &c, int chi, const u64 *ct, int n, u64 *plain) {
    u64 *x = c.ws_alloc((size_t)n * c.k * c.N);
    dot_with_secret(c, chi, ct, n, x);
    c.check(launch_decrypt_round(x, plain, n, c.k, c.logN, c.d_bc, c.ch[chi].pc, c.stream), "decrypt_round");
    c.op_count[Context::OP_DECRYPT] += (uint64_t)n;
}
// Decryptor::invariant_noise_budget: bits(q) - bits(|t * (c0 + c1 s) mod q|_centred) - 1, composed on the host
int op_noise_budget(Context &c, int chi, const u64 *ct) {
    const int k = c.k;
    const size_t N = c.N;
    u64 *x = c.ws_alloc((size_t)k * N);
    dot_with_secret(c, chi, ct, 1, x);
    std::vector<u64> h((size_t)k * N);
    CNHE_CUDA(cudaMemcpyAsync(h.data(), x, h.size() * 8, cudaMemcpyDeviceToHost, c.stream));
    c.sync();
    typedef std::vector<u64> Big;
    auto mul_small = [](const Big &a, u64 b) {
        Big r(a.size() + 1, 0);
        u64 carry = 0;
        for (size_t i = 0; i < a.size(); i++) {
            unsigned __int128 v = (unsigned __int128)a[i] * b + carry;
            r[i] = (u64)v;
            carry = (u64)(v >> 64);
        }
        r[a.size()] = carry;
        return r;
    };
    auto add_to = [](Big &a, const Big &b) {
        if (a.size() < b.size() + 1) a.resize(b.size() + 1, 0);
        u64 carry = 0;
        for (size_t i = 0; i < a.size(); i++) {
            unsigned __int128 v = (unsigned __int128)a[i] + (i < b.size() ? b[i] : 0) + carry;
            a[i] = (u64)v;
            carry = (u64)(v >> 64);
        }
    };
    auto cmp = [](const Big &a, const Big &b) {
        size_t n = std::max(a.size(), b.size());
        for (size_t i = n; i-- > 0;) {
            u64 x = i < a.size() ? a[i] : 0, y = i < b.size() ? b[i] : 0;
            if (x != y) return x < y ? -1 : 1;
        }
        return 0;
    };
    auto sub_from = [](Big &a, const Big &b) {
        u64 borrow = 0;
        for (size_t i = 0; i < a.size(); i++) {
            u64 y = i < b.size() ? b[i] : 0;
            unsigned __int128 v = (unsigned __int128)a[i] - y - borrow;
            a[i] = (u64)v;
            borrow = (u64)(v >> 64) ? 1 : 0;
        }
    };
    auto bits = [](const Big &a) {
        for (size_t i = a.size(); i-- > 0;)
            if (a[i]) return (int)(i * 64 + 64 - __builtin_clzll(a[i]));
        return 0;
    };
    Big Q{1};
    for (u64 p : c.q) Q = mul_small(Q, p);
    std::vector<Big> qhat(k, Big{1});
    for (int i = 0; i < k; i++)
        for (int l = 0; l < k; l++)
            if (l != i) qhat[i] = mul_small(qhat[i], c.q[l]);
    Big half = Q;
    {
        unsigned __int128 r = 0;
        for (size_t i = half.size(); i-- > 0;) {
            unsigned __int128 cur = (r << 64) | half[i];
            half[i] = (u64)(cur / 2);
            r = cur % 2;
        }
    }
    int maxbits = 0;
    const u64 t = c.ch[chi].t;
    for (size_t n = 0; n < N; n++) {
        Big acc{0};
        for (int i = 0; i < k; i++) {
            u64 v = hm::mul(hm::mul(h[i * N + n], t % c.q[i], c.q[i]), c.h_bc.inv_qhat_mod_q[i], c.q[i]);
            add_to(acc, mul_small(qhat[i], v));
        }
        while (cmp(acc, Q) >= 0) sub_from(acc, Q);
        if (cmp(acc, half) > 0) {
            Big tmp = Q;
            tmp.resize(std::max(tmp.size(), acc.size()), 0);
            sub_from(tmp, acc);
            acc = tmp;
        }
        maxbits = std::max(maxbits, bits(acc));
    }
    int b = bits(Q) - maxbits - 1;
    return b < 0 ? 0 : b;
}

// ---------------------------------------------------------------- keys (KeyGenerator of SEAL 3.2, sampled on the device)
BufRef &key_slot(Context &c, int channel, int what, u64 arg, size_t &words, bool create) {
    if (channel < 0 || channel >= c.P) throw Error(-1, "bad channel");
    Channel &ch = c.ch[channel];
    const size_t kN = (size_t)c.k * c.N;
    BufRef *slot = nullptr;
    switch (what) {
    case 0: words = kN; slot = &ch.sk; break;
    case 1: words = 2 * kN; slot = &ch.pk; break;
    case 2: words = (size_t)c.dm_relin.D * 2 * kN; slot = &ch.rlk; break;
    case 3:
        words = (size_t)c.dm_galois.D * 2 * kN;
        if (!create && !ch.glk.count(arg)) throw Error(-3, "Galois key not present");
        slot = &ch.glk[arg];
        break;
    default: throw Error(-1, "bad key kind");
    }
    if (create && !*slot) *slot = c.alloc(words);
    if (!*slot) throw Error(-3, "key is missing");
    return *slot;
}
// key-switching keys for `target` (k*N, NTT form): key (i,j) = (-(a s + e) + [residue i] 2^{jw} target, a)
static void make_kskeys(Context &c, Channel &ch, const u64 *target_ntt, const DigitMap &dm, int w, u64 purpose_a, u64 purpose_e, u64 key_tag, u64 *out) {
    const int k = c.k, D = dm.D;
    const size_t N = c.N, kN = (size_t)k * N;
    // c1 = a (uniform), written in place: key d part 1
    u64 *e = c.ws_alloc((size_t)D * kN), *as = c.ws_alloc((size_t)D * kN), *a = c.ws_alloc((size_t)D * kN);
    c.check(launch_sample(a, D, SAMPLE_UNIFORM, ch.rng, stream_id(purpose_a, key_tag * 256, 0), 1ULL << 16, k, c.logN, c.d_bc, c.stream), "sample");
    c.check(launch_sample(e, D, SAMPLE_NOISE, ch.rng, stream_id(purpose_e, key_tag * 256, 0), 1ULL << 16, k, c.logN, c.d_bc, c.stream), "sample");
    c.check(launch_ntt_forward(e, e, D * k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
    c.check(launch_dyadic_bcast(a, ch.sk->p, as, D, 1, 1, 0, k, c.logN, c.d_bc, c.stream), "dyadic");
    c.check(launch_ct_add(as, e, as, (size_t)D * kN, k, c.logN, c.d_bc, 0, c.stream), "ct_add");
    c.check(launch_ct_negate(as, as, (size_t)D * kN, k, c.logN, c.d_bc, c.stream), "ct_negate");
    // interleave into [D][2][k][N]
    CNHE_CUDA(cudaMemcpy2DAsync(out, 2 * kN * 8, as, kN * 8, kN * 8, D, cudaMemcpyDeviceToDevice, c.stream));
    CNHE_CUDA(cudaMemcpy2DAsync(out + kN, 2 * kN * 8, a, kN * 8, kN * 8, D, cudaMemcpyDeviceToDevice, c.stream));
    // add 2^{jw} * target on residue src[d] of c0 of key d
    std::vector<u64> factors(D);
    for (int d = 0; d < D; d++) factors[d] = hm::pw(2, (u64)dm.shift[d], c.q[dm.src[d]]);
    (void)w;
    u64 *dfac = c.ws_alloc(D);
    c.h2d(dfac, factors.data(), D * 8);
    c.check(launch_key_add_scaled(out, target_ntt, dfac, dm, k, c.logN, c.d_bc, c.stream), "key_add_scaled");
}
static void keys_generate_impl(Context &c, bool secure, u64 seed) {
    const int k = c.k;
    const size_t N = c.N, kN = (size_t)k * N;
    for (int ci = 0; ci < c.P; ci++) {
        c.set_channel(ci);
        Channel &ch = c.ch[ci];
        if (secure) rng_from_os(ch.rng);
        else { memset(&ch.rng, 0, sizeof(ch.rng)); ch.rng.seed = seed + (u64)ci; }
        ch.nonce = 1;
        size_t words;
        // secret key: ternary, kept in NTT form (and a coefficient-form copy for the Galois keys)
        BufRef &sk = key_slot(c, ci, 0, 0, words, true);
        u64 *sk_coeff = c.ws_alloc(kN);
        c.check(launch_sample(sk_coeff, 1, SAMPLE_TERNARY, ch.rng, stream_id(1, 0, 0), 0, k, c.logN, c.d_bc, c.stream), "sample");
        c.check(launch_ntt_forward(sk_coeff, sk->p, k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
        ch.have_sk = true;
        // public key (-(a s + e), a)
        BufRef &pk = key_slot(c, ci, 1, 0, words, true);
        u64 *e = c.ws_alloc(kN), *as = c.ws_alloc(kN);
        c.check(launch_sample(pk->p + kN, 1, SAMPLE_UNIFORM, ch.rng, stream_id(2, 0, 0), 0, k, c.logN, c.d_bc, c.stream), "sample");
        c.check(launch_sample(e, 1, SAMPLE_NOISE, ch.rng, stream_id(3, 0, 0), 0, k, c.logN, c.d_bc, c.stream), "sample");
        c.check(launch_ntt_forward(e, e, k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
        c.check(launch_dyadic_bcast(pk->p + kN, sk->p, as, 1, 1, 1, 0, k, c.logN, c.d_bc, c.stream), "dyadic");
        c.check(launch_ct_add(as, e, as, kN, k, c.logN, c.d_bc, 0, c.stream), "ct_add");
        c.check(launch_ct_negate(as, pk->p, kN, k, c.logN, c.d_bc, c.stream), "ct_negate");
        ch.have_pk = true;
        // relinearization keys for s^2
        BufRef &rlk = key_slot(c, ci, 2, 0, words, true);
        u64 *s2 = c.ws_alloc(kN);
        c.check(launch_dyadic_bcast(sk->p, sk->p, s2, 1, 1, 1, 0, k, c.logN, c.d_bc, c.stream), "dyadic");
        make_kskeys(c, ch, s2, c.dm_relin, c.dbc_relin, 4, 5, 0, rlk->p);
        ch.have_rlk = true;
        // Galois keys: s(x^elt) in NTT form
        for (size_t gi = 0; gi < c.galois_elts.size(); gi++) {
            const u64 elt = c.galois_elts[gi], m2 = 2ULL * N;
            u64 einv = 0;
            for (u64 x = 1; x < m2; x += 2)
                if (((x * elt) & (m2 - 1)) == 1) { einv = x; break; }
            // reuse the ciphertext Galois kernel on (sk_coeff, sk_coeff): perm_c1 receives the permuted second part
            u64 *pair = c.ws_alloc(2 * kN), *base = c.ws_alloc(2 * kN), *rs = c.ws_alloc(kN);
            CNHE_CUDA(cudaMemcpyAsync(pair, sk_coeff, kN * 8, cudaMemcpyDeviceToDevice, c.stream));
            CNHE_CUDA(cudaMemcpyAsync(pair + kN, sk_coeff, kN * 8, cudaMemcpyDeviceToDevice, c.stream));
            c.check(launch_galois(pair, base, rs, 1, einv, k, c.logN, c.d_bc, c.stream), "galois");
            c.check(launch_ntt_forward(rs, rs, k, c.logN, c.d_tabs, 0, k, fp_range(c, 0, k), c.stream), "ntt_forward");
            BufRef &gk = key_slot(c, ci, 3, elt, words, true);
            make_kskeys(c, ch, rs, c.dm_galois, c.dbc_galois, 6, 7, gi + 1, gk->p);
        }
        c.sync();
        ws_release_all(c);
    }
}

void keys_generate(Context &c, u64 seed) { keys_generate_impl(c, false, seed); }
void keys_generate_secure(Context &c) { keys_generate_impl(c, true, 0); }

} // namespace cnhe
