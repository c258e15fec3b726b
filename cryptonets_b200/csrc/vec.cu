// C ABI of libcnhe (include/cnhe.h) and the vector layer behind it.
//
// One cnhe_vec is the reference's EncryptedSealBfvVector ("HE Wrapper/EncryptedSealBfvVector.cs:150-573"): P channels,
// one per plaintext modulus, each carrying what AtomicSealBfvEncryptedVector ("HE Wrapper/AtomicSealBfvVector.cs:303-326")
// keeps in `Ciphertext[] encData` / `Plaintext[] plainData` -- here blocks of HBM.  The functions below restate that
// class method by method (cited inline); the arithmetic itself is in the CUDA kernels.
#include <cmath>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <unordered_map>

#include "hostmath.h"
#include "vec.h"

typedef unsigned __int128 u128;

static thread_local std::string g_err;
int set_err(int code, const std::string &m) {
    g_err = m;
    return code;
}
extern "C" const char *cnhe_last_error(void) { return g_err.c_str(); }
extern "C" const char *cnhe_version(void) { return "cnhe-b200 0.2 (sm_100a)"; }

// ---------------------------------------------------------------------------------------------------- context & keys
extern "C" int cnhe_context_create_custom(const uint64_t *plain_primes, int P, uint32_t N, const uint64_t *coeff, int k, int dbc_relin,
                                          int dbc_galois, int device, cnhe_ctx **out) {
    try {
        if (!plain_primes || !coeff || !out) fail("null argument");
        std::vector<u64> pp(plain_primes, plain_primes + P), cc(coeff, coeff + k);
        Context *c = context_create(pp.data(), P, N, cc.data(), k, dbc_relin, dbc_galois, device);
        *out = new cnhe_ctx{c};
    } catch (const Error &e) { return set_err(e.code, e.what()); } catch (const std::exception &e) { return set_err(CNHE_ERR_INVALID, e.what()); }
    return CNHE_OK;
}
extern "C" int cnhe_context_create(const uint64_t *plain_primes, int P, uint32_t N, int dbc_relin, int dbc_galois, int small_modulus_count,
                                   int device, cnhe_ctx **out) {
    std::vector<u64> q = default_coeff_modulus(N);
    if (q.empty()) return set_err(CNHE_ERR_INVALID, "no default coefficient modulus for this PolyModulusDegree");
    if (small_modulus_count > 0 && small_modulus_count < (int)q.size()) q.resize(small_modulus_count); // AtomicSealBfvVector.cs:148-149
    std::vector<uint64_t> qq(q.begin(), q.end());
    return cnhe_context_create_custom(plain_primes, P, N, qq.data(), (int)qq.size(), dbc_relin, dbc_galois, device, out);
}
extern "C" int cnhe_context_destroy(cnhe_ctx *h) {
    if (!h) return CNHE_OK;
    delete h->c;
    delete h;
    return CNHE_OK;
}
extern "C" int cnhe_context_info(const cnhe_ctx *h, uint32_t *N, int *k, int *P, int *relin_digits, int *galois_digits, int *galois_elts) {
    if (!h) return set_err(CNHE_ERR_INVALID, "null context");
    const Context &c = *h->c;
    if (N) *N = c.N;
    if (k) *k = c.k;
    if (P) *P = c.P;
    if (relin_digits) *relin_digits = c.dm_relin.D;
    if (galois_digits) *galois_digits = c.dm_galois.D;
    if (galois_elts) *galois_elts = (int)c.galois_elts.size();
    return CNHE_OK;
}
extern "C" int cnhe_context_coeff_moduli(const cnhe_ctx *h, uint64_t *out) {
    if (!h || !out) return set_err(CNHE_ERR_INVALID, "null argument");
    for (int i = 0; i < h->c->k; i++) out[i] = h->c->q[i];
    return CNHE_OK;
}
extern "C" int cnhe_context_bsk_moduli(const cnhe_ctx *h, uint64_t *out, int *count) {
    if (!h || !count) return set_err(CNHE_ERR_INVALID, "null argument");
    *count = h->c->kb;
    if (out)
        for (int i = 0; i < h->c->kb; i++) out[i] = h->c->bsk[i];
    return CNHE_OK;
}
extern "C" int cnhe_context_plain_moduli(const cnhe_ctx *h, uint64_t *out) {
    if (!h || !out) return set_err(CNHE_ERR_INVALID, "null argument");
    for (int i = 0; i < h->c->P; i++) out[i] = h->c->t[i];
    return CNHE_OK;
}
extern "C" int cnhe_context_galois_elts(const cnhe_ctx *h, uint64_t *out) {
    if (!h || !out) return set_err(CNHE_ERR_INVALID, "null argument");
    for (size_t i = 0; i < h->c->galois_elts.size(); i++) out[i] = h->c->galois_elts[i];
    return CNHE_OK;
}
extern "C" int cnhe_context_set_option(cnhe_ctx *h, const char *name, int64_t value) {
    API_BEGIN(h)
    std::string n(name ? name : "");
    if (n == "behz_centered_mtilde") {
        c.h_bc.centered_mtilde = value ? 1 : 0;
        c.h_bf.centered_mtilde = value ? 1 : 0;
        CNHE_CUDA(cudaMemcpyAsync(c.d_bc, &c.h_bc, sizeof(BehzConst), cudaMemcpyHostToDevice, c.stream));
        CNHE_CUDA(cudaMemcpyAsync(c.d_bf, &c.h_bf, sizeof(BehzConstF), cudaMemcpyHostToDevice, c.stream));
        c.sync();
    } else if (n == "multi_stream") {
        // buffers remember the stream they are released on: switching the stream mode with work or uploads in flight would let the
        // recycler hand a block out while another stream still reads it.  Quiesce, drop the per-stream free lists, then switch.
        for (const Context::UploadSlot &u : c.upload_slots)
            if (u.busy) fail("multi_stream cannot change while imported batches are alive (dispose them first)");
        c.sync();
        CNHE_CUDA(cudaStreamSynchronize(c.copy_stream));
        c.drop_recycled();
        c.multi_stream = value != 0;
    } else if (n == "trace_noise") {
        c.trace_noise = value != 0;
        c.trace.clear();
        if (!c.trace_noise) c.budget_of.clear();
    } else if (n == "chunk") {
        if (value < 1 || value > 4096) fail("chunk must be in [1,4096]");
        c.chunk = (int)value;
    } else fail("unknown option");
    API_END
}
// Interop with the caller's own GPU work (NCCL collectives, torch copies): the CUDA stream of a plaintext-modulus channel, and the two
// fences of the per-channel streams -- join: stream 0 waits for the tail of every channel; fork: every channel waits for stream 0.
// A caller that enqueues on stream 0 between a join and a fork is ordered after everything queued so far and before everything queued later.
extern "C" int cnhe_context_stream(cnhe_ctx *h, int channel, uint64_t *stream) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P || !stream) fail("bad arguments");
    *stream = (uint64_t)c.streams[c.multi_stream ? channel : 0];
    API_END
}
extern "C" int cnhe_context_join_streams(cnhe_ctx *h) {
    API_BEGIN(h)
    c.join_streams();
    API_END
}
extern "C" int cnhe_context_fork_streams(cnhe_ctx *h) {
    API_BEGIN(h)
    c.fork_streams();
    API_END
}
extern "C" int cnhe_context_sync(cnhe_ctx *h) {
    API_BEGIN(h)
    c.sync();
    API_END
}
extern "C" uint64_t cnhe_kernel_launch_count(const cnhe_ctx *h) { return h ? h->c->launches : 0; }
extern "C" int cnhe_keys_generate(cnhe_ctx *h, uint64_t seed) {
    API_BEGIN(h)
    keys_generate(c, seed);
    API_END
}
extern "C" int cnhe_keys_generate_secure(cnhe_ctx *h) {
    API_BEGIN(h)
    keys_generate_secure(c);
    API_END
}
// OperationsCount / CryptoTracker mirrors
extern "C" int cnhe_op_counts(cnhe_ctx *h, uint64_t *out, int cap, int reset) {
    API_BEGIN(h)
    if (!out || cap < Context::OP_COUNT) fail("need room for CNHE_OP_COUNT counters");
    for (int i = 0; i < Context::OP_COUNT; i++) out[i] = c.op_count[i];
    if (reset) for (int i = 0; i < Context::OP_COUNT; i++) c.op_count[i] = 0;
    API_END
}
extern "C" const char *cnhe_op_name(int kind) { return op_kind_name(kind); }
extern "C" int cnhe_trace_read(cnhe_ctx *h, int32_t *out, size_t cap_records, size_t *n_records, int clear) {
    API_BEGIN(h)
    if (n_records) *n_records = c.trace.size();
    if (out) {
        const size_t n = std::min(cap_records, c.trace.size());
        static_assert(sizeof(Context::TraceRec) == 8 * sizeof(int32_t), "trace record layout");
        memcpy(out, c.trace.data(), n * sizeof(Context::TraceRec));
    }
    if (clear) { c.trace.clear(); }
    API_END
}
extern "C" int cnhe_keys_set_seed(cnhe_ctx *h, int channel, uint64_t seed) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P) fail("bad channel");
    memset(&c.ch[channel].rng, 0, sizeof(RngKey)); // deterministic sampler (tests): see cnhe.h
    c.ch[channel].rng.seed = seed;
    API_END
}
extern "C" int cnhe_keys_export(cnhe_ctx *h, int channel, int what, uint64_t arg, uint64_t *dst, size_t cap) {
    API_BEGIN(h)
    size_t words;
    BufRef &b = key_slot(c, channel, what, arg, words, false);
    if (cap < words) fail("destination too small");
    CNHE_CUDA(cudaMemcpyAsync(dst, b->p, words * 8, cudaMemcpyDeviceToHost, c.stream));
    c.sync();
    API_END
}
extern "C" int cnhe_keys_import(cnhe_ctx *h, int channel, int what, uint64_t arg, const uint64_t *src, size_t nwords) {
    API_BEGIN(h)
    size_t words;
    BufRef &b = key_slot(c, channel, what, arg, words, true);
    if (nwords != words) fail("wrong key size");
    CNHE_CUDA(cudaMemcpyAsync(b->p, src, words * 8, cudaMemcpyHostToDevice, c.stream));
    c.sync();
    Channel &ch = c.ch[channel];
    if (what == 0) ch.have_sk = true;
    if (what == 1) ch.have_pk = true;
    if (what == 2) ch.have_rlk = true;
    API_END
}

// ---------------------------------------------------------------------------------------------------- raw / microbench
extern "C" int cnhe_dev_alloc(cnhe_ctx *h, size_t words, uint64_t *dptr) {
    API_BEGIN(h)
    void *p = nullptr;
    CNHE_CUDA(cudaMalloc(&p, words * 8));
    *dptr = (uint64_t)p;
    API_END
}
extern "C" int cnhe_dev_free(cnhe_ctx *h, uint64_t dptr) {
    API_BEGIN(h)
    c.sync();
    CNHE_CUDA(cudaFree((void *)dptr));
    API_END
}
extern "C" int cnhe_dev_upload(cnhe_ctx *h, uint64_t dptr, const uint64_t *src, size_t words) {
    API_BEGIN(h)
    CNHE_CUDA(cudaMemcpyAsync((void *)dptr, src, words * 8, cudaMemcpyHostToDevice, c.stream));
    c.sync();
    API_END
}
extern "C" int cnhe_dev_download(cnhe_ctx *h, uint64_t *dst, uint64_t dptr, size_t words) {
    API_BEGIN(h)
    CNHE_CUDA(cudaMemcpyAsync(dst, (void *)dptr, words * 8, cudaMemcpyDeviceToHost, c.stream));
    c.sync();
    API_END
}
extern "C" int cnhe_raw_ntt(cnhe_ctx *h, uint64_t src, uint64_t dst, int n_polys, int mod_base, int mod_count, int inverse) {
    API_BEGIN(h)
    if (mod_base < 0 || mod_count < 1 || mod_base + mod_count > c.k + c.kb + c.P) fail("bad modulus range");
    op_ntt(c, (const u64 *)src, (u64 *)dst, n_polys, mod_base, mod_count, inverse != 0);
    API_END
}
static std::vector<const u64 *> strided(uint64_t base, int n, size_t words) {
    std::vector<const u64 *> v(n);
    for (int i = 0; i < n; i++) v[i] = (const u64 *)base + (size_t)i * words;
    return v;
}
extern "C" int cnhe_raw_multiply(cnhe_ctx *h, int channel, uint64_t a, uint64_t b, int n, uint64_t out3) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P) fail("bad channel");
    c.set_channel(channel);
    op_multiply(c, channel, strided(a, n, c.ct_words()), strided(b, n, c.ct_words()), (u64 *)out3);
    API_END
}
extern "C" int cnhe_raw_relinearize(cnhe_ctx *h, int channel, uint64_t in3, int n, uint64_t out2) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P) fail("bad channel");
    c.set_channel(channel);
    op_relinearize(c, channel, (const u64 *)in3, n, (u64 *)out2);
    API_END
}
extern "C" int cnhe_raw_multiply_relin(cnhe_ctx *h, int channel, uint64_t a, uint64_t b, int n, uint64_t out2) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P) fail("bad channel");
    c.set_channel(channel);
    op_multiply_relin(c, channel, strided(a, n, c.ct_words()), strided(b, n, c.ct_words()), (u64 *)out2);
    API_END
}
extern "C" int cnhe_raw_apply_galois(cnhe_ctx *h, int channel, uint64_t in, int n, uint64_t elt, uint64_t out) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P) fail("bad channel");
    c.set_channel(channel);
    op_apply_galois(c, channel, (const u64 *)in, n, elt, (u64 *)out);
    API_END
}
extern "C" int cnhe_raw_rotate_rows(cnhe_ctx *h, int channel, uint64_t in, int n, int steps, uint64_t out) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P) fail("bad channel");
    c.set_channel(channel);
    op_rotate_rows(c, channel, (const u64 *)in, n, steps, (u64 *)out);
    API_END
}
extern "C" int cnhe_raw_behz_lift(cnhe_ctx *h, uint64_t in_cts, int n, uint64_t out) {
    API_BEGIN(h)
    if (c.fp_elementwise)
        c.check(launch_behz_lift_fp(upload_ptrs(c, strided(in_cts, n, c.ct_words())), (u64 *)out, n, c.logN, &c.h_bf, 0, c.stream), "behz_lift_fp");
    else c.check(launch_behz_lift(upload_ptrs(c, strided(in_cts, n, c.ct_words())), (u64 *)out, n, c.logN, c.d_bc, c.stream), "behz_lift");
    API_END
}
extern "C" int cnhe_raw_behz_floor(cnhe_ctx *h, int channel, uint64_t d, int n, uint64_t out3) {
    API_BEGIN(h)
    if (channel < 0 || channel >= c.P) fail("bad channel");
    c.set_channel(channel);
    if (c.fp_elementwise) c.check(launch_behz_floor_fp((const u64 *)d, (u64 *)out3, n, c.ch[channel].t, c.logN, &c.h_bf, 0, c.stream), "behz_floor_fp");
    else c.check(launch_behz_floor((const u64 *)d, (u64 *)out3, n, c.ch[channel].t, c.logN, c.d_bc, c.stream), "behz_floor");
    API_END
}
extern "C" int cnhe_dev_copy(cnhe_ctx *h, uint64_t dst, uint64_t src, size_t words) {
    API_BEGIN(h)
    // the pointers may belong to any channel: order the copy after every channel's queued work and before anything queued later
    c.join_streams();
    CNHE_CUDA(cudaMemcpyAsync((void *)dst, (const void *)src, words * 8, cudaMemcpyDeviceToDevice, c.streams[0]));
    c.fork_streams();
    API_END
}
extern "C" int cnhe_prof_enable(cnhe_ctx *h, int on) {
    API_BEGIN(h)
    c.prof_flush();
    c.prof = on != 0;
    if (on) for (int i = 0; i < 6; i++) { c.prof_ms[i] = 0; c.prof_bytes[i] = 0; c.prof_n[i] = 0; }
    API_END
}
extern "C" int cnhe_prof_collect(cnhe_ctx *h, int family, double *total_ms, uint64_t *launches, double *bytes) {
    API_BEGIN(h)
    if (family < 0 || family > 5) fail("bad family");
    c.prof_flush();
    if (total_ms) *total_ms = c.prof_ms[family];
    if (launches) *launches = c.prof_n[family];
    if (bytes) *bytes = c.prof_bytes[family];
    API_END
}
extern "C" int cnhe_raw_event_timing(cnhe_ctx *h, int start) {
    API_BEGIN(h)
    if (start) {
        c.join_streams();
        CNHE_CUDA(cudaEventRecord(c.ev0, c.streams[0]));
        c.fork_streams();
    } else {
        c.join_streams();
        CNHE_CUDA(cudaEventRecord(c.ev1, c.streams[0]));
    }
    API_END
}
extern "C" int cnhe_raw_elapsed_ms(cnhe_ctx *h, float *ms) {
    API_BEGIN(h)
    CNHE_CUDA(cudaEventSynchronize(c.ev1));
    CNHE_CUDA(cudaEventElapsedTime(ms, c.ev0, c.ev1));
    API_END
}

// ---------------------------------------------------------------------------------------------------- vector helpers
cnhe_vec *new_vec(Context &c, uint64_t dim, double scale, int format, bool enc, int blocks) {
    cnhe_vec *v = new cnhe_vec();
    v->ctx = &c;
    v->dim = dim;
    v->scale = scale;
    v->format = format;
    v->enc = enc;
    v->blocks = blocks;
    v->buf.resize(c.P);
    v->off.assign(c.P, 0);
    return v;
}
void alloc_channels(cnhe_vec *v) {
    for (int ch = 0; ch < v->ctx->P; ch++) {
        v->ctx->set_channel(ch);
        v->buf[ch] = v->ctx->alloc((size_t)v->blocks * v->unit());
    }
}
static cnhe_vec *alias_of(const cnhe_vec *a) { return new cnhe_vec(*a); } // shares the reference-counted buffers
void same_ctx(Context &c, const cnhe_vec *v) {
    if (!v) fail("null vector");
    if (v->ctx != &c) fail("vector belongs to another context");
}
// SplitBigNumbers ("EncryptedSealBfvVector.cs:352-365"): round(v*scale) -> +bigFactor if negative -> residues
static void split_values(Context &c, const double *v, uint64_t n, double scale, std::vector<std::vector<u64>> &res) {
    res.assign(c.P, std::vector<u64>(n));
    for (uint64_t j = 0; j < n; j++) {
        const double w = std::nearbyint(v[j] * scale); // Math.Round: to nearest, ties to even
        if (!(std::fabs(w) < 1.5e38)) fail("value out of range");
        const bool neg = w < 0;
        u128 mag = (u128)(neg ? -w : w);
        if (mag >= c.big_factor) mag %= c.big_factor;
        const u128 z = (neg && mag) ? c.big_factor - mag : mag;
        for (int i = 0; i < c.P; i++) res[i][j] = (u64)(z % c.t[i]);
    }
}
static u128 mulmod_u128(u128 a, u64 b, u128 m) { // a < m < 2^126
    u128 r = 0;
    while (b) {
        if (b & 1) { r += a; if (r >= m) r -= m; }
        a += a; if (a >= m) a -= m;
        b >>= 1;
    }
    return r;
}
// JoinSplitNumbers ("EncryptedSealBfvVector.cs:381-395")
static void join_values(Context &c, const std::vector<std::vector<u64>> &split, uint64_t n, double scale, double *out) {
    for (uint64_t j = 0; j < n; j++) {
        u128 acc = 0;
        for (int i = 0; i < c.P; i++) {
            acc += mulmod_u128(c.crt_coeff[i] % c.big_factor, split[i][j], c.big_factor);
            if (acc >= c.big_factor) acc -= c.big_factor;
        }
        double val;
        if (acc * 2 > c.big_factor) val = -(double)(c.big_factor - acc);
        else val = (double)acc;
        out[j] = val / scale;
    }
}

static cnhe_vec *make_vector_split(Context &c, const std::vector<std::vector<u64>> &split, const double *v, uint64_t dim, double scale, int format,
                                   bool encrypt);
static cnhe_vec *make_vector(Context &c, const double *v, uint64_t dim, double scale, int format, bool encrypt) {
    if (!v && dim) fail("null values");
    if (format != CNHE_DENSE && format != CNHE_SPARSE) fail("bad format");
    if (scale == 0) scale = 1; // AtomicSealBfvVector.cs:1120
    std::vector<std::vector<u64>> split;
    split_values(c, v, dim, scale, split);
    return make_vector_split(c, split, v, dim, scale, format, encrypt);
}
// `split`: per plaintext modulus the residues of the (already scaled and rounded) values; `v` (may be null) only feeds the
// all-slots-equal shortcut of plain dense vectors
static cnhe_vec *make_vector_split(Context &c, const std::vector<std::vector<u64>> &split, const double *v, uint64_t dim, double scale, int format,
                                   bool encrypt) {
    const size_t N = c.N;
    const int blocks = format == CNHE_DENSE ? (int)((dim + N - 1) / N) : (int)dim;
    if (blocks < 1) fail("empty vector");
    cnhe_vec *out = new_vec(c, dim, scale, format, encrypt, blocks);
    std::unique_ptr<cnhe_vec> guard(out);
    if (format == CNHE_SPARSE && !encrypt) {
        out->scalars = split;
        for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
            out->buf[ch] = c.alloc(dim);
            CNHE_CUDA(cudaMemcpyAsync(out->buf[ch]->p, split[ch].data(), dim * 8, cudaMemcpyHostToDevice, c.stream));
        }
        c.sync();
        return guard.release();
    }
    alloc_channels(out);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        if (format == CNHE_DENSE) {
            // BatchEncoder.Encode per N-slot chunk (AtomicSealBfvVector.cs:1123-1133); a short last chunk is zero padded
            std::vector<u64> padded((size_t)blocks * N, 0);
            memcpy(padded.data(), split[ch].data(), dim * 8);
            u64 *dvals = c.ws_alloc((size_t)blocks * N);
            CNHE_CUDA(cudaMemcpyAsync(dvals, padded.data(), padded.size() * 8, cudaMemcpyHostToDevice, c.stream));
            u64 *plain = encrypt ? c.ws_alloc((size_t)blocks * N) : out->ptr(ch);
            op_encode(c, ch, dvals, blocks, (int)N, plain);
            if (encrypt) {
                op_encrypt(c, ch, plain, N, blocks, (int)N, take_nonces(c, ch, blocks), out->ptr(ch));
            }
        } else { // sparse encrypted: one constant-polynomial plaintext per element (AtomicSealBfvVector.cs:1135-1138)
            u64 *dvals = c.ws_alloc(dim);
            CNHE_CUDA(cudaMemcpyAsync(dvals, split[ch].data(), dim * 8, cudaMemcpyHostToDevice, c.stream));
            op_encrypt(c, ch, dvals, 1, blocks, 1, take_nonces(c, ch, blocks), out->ptr(ch));
        }
        c.sync(); // host staging buffers go out of scope
    }
    if (v && !encrypt && format == CNHE_DENSE && dim % N == 0) { // every slot equal => every plaintext is the constant polynomial
        bool all_eq = true;
        for (uint64_t j = 1; j < dim && all_eq; j++) all_eq = v[j] == v[0];
        if (all_eq) {
            out->is_const = true;
            for (int ch = 0; ch < c.P; ch++) out->const_val.push_back(split[ch][0]);
        }
    }
    return guard.release();
}

extern "C" int cnhe_vec_encrypt(cnhe_ctx *h, const double *v, uint64_t dim, double scale, int format, cnhe_vec **out) {
    API_BEGIN(h)
    *out = make_vector(c, v, dim, scale, format, true);
    API_END
}
// BigInteger entry points of the factory ("HE Wrapper/IFactory.cs:29,43" -> EncryptedSealBfvVector(IEnumerable<BigInteger>, ...),
// "EncryptedSealBfvVector.cs:188-199"): the host reduces each big integer modulo every plaintext prime (SplitBigNumbers, ":367-379")
extern "C" int cnhe_vec_from_residues(cnhe_ctx *h, const uint64_t *residues, uint64_t dim, double scale, int format, int encrypt, cnhe_vec **out) {
    API_BEGIN(h)
    if (!residues || !out || dim < 1) fail("bad arguments");
    if (format != CNHE_DENSE && format != CNHE_SPARSE) fail("bad format");
    std::vector<std::vector<u64>> split(c.P, std::vector<u64>(dim));
    for (int ch = 0; ch < c.P; ch++)
        for (uint64_t j = 0; j < dim; j++) {
            split[ch][j] = residues[(size_t)ch * dim + j];
            if (split[ch][j] >= c.t[ch]) fail("residue not reduced modulo its plaintext prime");
        }
    *out = make_vector_split(c, split, nullptr, dim, scale == 0 ? 1 : scale, format, encrypt != 0);
    API_END
}
extern "C" int cnhe_vec_plain(cnhe_ctx *h, const double *v, uint64_t dim, double scale, int format, cnhe_vec **out) {
    API_BEGIN(h)
    *out = make_vector(c, v, dim, scale, format, false);
    API_END
}
// n dense single-block-or-more vectors encrypted in one wave per channel (GetEncryptedMatrix, IFactory.cs:353-380)
extern "C" int cnhe_vecs_encrypt(cnhe_ctx *h, const double *v, int n, uint64_t dim, double scale, cnhe_vec **out) {
    API_BEGIN(h)
    if (n < 1 || !v || !out) fail("bad arguments");
    if (scale == 0) scale = 1;
    const size_t N = c.N;
    const int bl = (int)((dim + N - 1) / N);
    std::vector<std::vector<u64>> split;
    split_values(c, v, (uint64_t)n * dim, scale, split);
    std::vector<BufRef> big(c.P);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        std::vector<u64> padded((size_t)n * bl * N, 0);
        for (int i = 0; i < n; i++) memcpy(&padded[(size_t)i * bl * N], &split[ch][(size_t)i * dim], dim * 8);
        u64 *dvals = c.ws_alloc(padded.size()), *plain = c.ws_alloc(padded.size());
        CNHE_CUDA(cudaMemcpyAsync(dvals, padded.data(), padded.size() * 8, cudaMemcpyHostToDevice, c.stream));
        op_encode(c, ch, dvals, n * bl, (int)N, plain);
        big[ch] = c.alloc((size_t)n * bl * c.ct_words());
        op_encrypt(c, ch, plain, N, n * bl, (int)N, take_nonces(c, ch, (u64)n * bl), big[ch]->p);
        c.sync();
    }
    for (int i = 0; i < n; i++) {
        cnhe_vec *o = new_vec(c, dim, scale, CNHE_DENSE, true, bl);
        for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
            o->buf[ch] = big[ch];
            o->off[ch] = (size_t)i * bl * c.ct_words();
        }
        out[i] = o;
    }
    API_END
}

// Decrypt of the blocks of one channel into residues (dense: first `dim` slots; sparse: constant coefficients)
static void decrypt_channel(Context &c, const cnhe_vec *v, int ch, std::vector<u64> &res) {
    const size_t N = c.N;
    res.assign(v->dim, 0);
    if (!v->enc && v->format == CNHE_SPARSE) { res = v->scalars[ch]; return; }
    u64 *plain;
    if (v->enc) {
        plain = c.ws_alloc((size_t)v->blocks * N);
        op_decrypt(c, ch, v->ptr(ch), v->blocks, plain);
    } else plain = v->ptr(ch);
    if (v->format == CNHE_DENSE) {
        u64 *vals = c.ws_alloc((size_t)v->blocks * N);
        op_decode(c, ch, plain, v->blocks, vals);
        std::vector<u64> hostv((size_t)v->blocks * N);
        CNHE_CUDA(cudaMemcpyAsync(hostv.data(), vals, hostv.size() * 8, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        const uint64_t take = std::min<uint64_t>(v->dim, hostv.size());
        memcpy(res.data(), hostv.data(), take * 8);
    } else {
        std::vector<u64> hostv(v->blocks);
        CNHE_CUDA(cudaMemcpy2DAsync(hostv.data(), 8, plain, N * 8, 8, v->blocks, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        const uint64_t take = std::min<uint64_t>(v->dim, hostv.size());
        memcpy(res.data(), hostv.data(), take * 8);
    }
}
extern "C" int cnhe_vec_decrypt(cnhe_ctx *h, const cnhe_vec *v, double *out, uint64_t cap) {
    API_BEGIN(h)
    same_ctx(c, v);
    if (cap < v->dim) fail("destination too small");
    std::vector<std::vector<u64>> split(c.P);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        decrypt_channel(c, v, ch, split[ch]);
    }
    join_values(c, split, v->dim, v->scale, out);
    API_END
}
// DecryptFullPrecision ("EncryptedSealBfvVector.cs:343-348"): the residues of every channel, for the caller's BigInteger CRT join
// (JoinSplitNumbers, ":397-411"); out is [P][dim]
extern "C" int cnhe_vec_decrypt_residues(cnhe_ctx *h, const cnhe_vec *v, uint64_t *out, uint64_t cap) {
    API_BEGIN(h)
    same_ctx(c, v);
    if (!out || cap < v->dim * (uint64_t)c.P) fail("destination too small");
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        std::vector<u64> r;
        decrypt_channel(c, v, ch, r);
        memcpy(out + (size_t)ch * v->dim, r.data(), v->dim * 8);
    }
    API_END
}
extern "C" int cnhe_vecs_decrypt(cnhe_ctx *h, const cnhe_vec *const *vecs, int n, double *out, uint64_t dim) {
    API_BEGIN(h)
    for (int i = 0; i < n; i++) {
        same_ctx(c, vecs[i]);
        if (vecs[i]->dim != dim) fail("all vectors must have the same dimension");
        std::vector<std::vector<u64>> split(c.P);
        for (int ch = 0; ch < c.P; ch++) {
            c.set_channel(ch);
            decrypt_channel(c, vecs[i], ch, split[ch]);
        }
        join_values(c, split, dim, vecs[i]->scale, out + (size_t)i * dim);
    }
    API_END
}
extern "C" int cnhe_vec_copy(cnhe_ctx *h, const cnhe_vec *v, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, v);
    cnhe_vec *o = new cnhe_vec(*v);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        const size_t words = (size_t)v->blocks * v->unit();
        o->buf[ch] = c.alloc(words);
        o->off[ch] = 0;
        CNHE_CUDA(cudaMemcpyAsync(o->buf[ch]->p, v->ptr(ch), words * 8, cudaMemcpyDeviceToDevice, c.stream));
    }
    *out = o;
    API_END
}
extern "C" int cnhe_vec_destroy(cnhe_vec *v) {
    if (!v) return CNHE_OK;
    try {
        std::lock_guard<std::recursive_mutex> lock(v->ctx->mu);
        cudaSetDevice(v->ctx->device);
        delete v;
    } catch (...) { return set_err(CNHE_ERR_INVALID, "destroy failed"); }
    return CNHE_OK;
}
extern "C" int cnhe_vecs_destroy(cnhe_vec *const *vecs, int n) {
    if (!vecs || n < 1) return CNHE_OK;
    try {
        Context *ctx = nullptr;
        for (int i = 0; i < n && !ctx; i++)
            if (vecs[i]) ctx = vecs[i]->ctx;
        if (!ctx) return CNHE_OK;
        std::lock_guard<std::recursive_mutex> lock(ctx->mu);
        cudaSetDevice(ctx->device);
        for (int i = 0; i < n; i++)
            if (vecs[i]) {
                if (vecs[i]->ctx != ctx) return set_err(CNHE_ERR_INVALID, "vectors of different contexts");
                delete vecs[i];
            }
    } catch (...) { return set_err(CNHE_ERR_INVALID, "destroy failed"); }
    return CNHE_OK;
}
extern "C" int cnhe_vec_meta(const cnhe_vec *v, uint64_t *dim, double *scale, int *format, int *is_encrypted, int *blocks, uint64_t *block_size) {
    if (!v) return set_err(CNHE_ERR_INVALID, "null vector");
    if (dim) *dim = v->dim;
    if (scale) *scale = v->scale;
    if (format) *format = v->format;
    if (is_encrypted) *is_encrypted = v->enc ? 1 : 0;
    if (blocks) *blocks = v->blocks;
    if (block_size) *block_size = v->ctx->N;
    return CNHE_OK;
}
extern "C" int cnhe_vec_register_scale(cnhe_vec *v, double scale) {
    if (!v) return set_err(CNHE_ERR_INVALID, "null vector");
    v->scale = scale;
    return CNHE_OK;
}
extern "C" int cnhe_vec_register_dim(cnhe_vec *v, uint64_t dim) {
    if (!v) return set_err(CNHE_ERR_INVALID, "null vector");
    v->dim = dim;
    return CNHE_OK;
}
extern "C" int cnhe_vec_export_raw(cnhe_ctx *h, const cnhe_vec *v, int channel, int block, uint64_t *dst, size_t cap) {
    API_BEGIN(h)
    same_ctx(c, v);
    if (!v->enc) fail("vector is not encrypted");
    if (channel < 0 || channel >= c.P || block < 0 || block >= v->blocks) fail("bad channel/block");
    if (cap < c.ct_words()) fail("destination too small");
    CNHE_CUDA(cudaMemcpyAsync(dst, v->block(channel, block), c.ct_words() * 8, cudaMemcpyDeviceToHost, c.stream));
    c.sync();
    API_END
}
extern "C" int cnhe_vec_import_raw(cnhe_ctx *h, const uint64_t *src, int blocks, uint64_t dim, double scale, int format, cnhe_vec **out) {
    API_BEGIN(h)
    if (!src || blocks < 1) fail("bad arguments");
    cnhe_vec *o = new_vec(c, dim, scale, format, true, blocks);
    alloc_channels(o);
    const size_t words = (size_t)blocks * c.ct_words();
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        CNHE_CUDA(cudaMemcpyAsync(o->ptr(ch), src + (size_t)ch * words, words * 8, cudaMemcpyDefault, c.stream)); // host or device source
    }
    c.sync();
    *out = o;
    API_END
}
extern "C" int cnhe_vecs_import_raw(cnhe_ctx *h, const uint64_t *src, int n, int blocks, uint64_t dim, double scale, int format, cnhe_vec **out) {
    API_BEGIN(h)
    if (!src || n < 1 || blocks < 1 || !out) fail("bad arguments");
    const size_t per = (size_t)blocks * c.ct_words(), words = (size_t)n * per;
    std::vector<BufRef> big(c.P);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        // the block comes from the context's rotating upload slots and the copy is ordered on the upload stream only (not behind the
        // kernels already queued on the channel's stream): channels follow each other over PCIe, channel 0 computes while channel 1 is
        // still uploading, and an import issued before the previous batch is exported overlaps that batch's kernels
        big[ch] = c.alloc_upload(words, c.stream);
        CNHE_CUDA(cudaMemcpyAsync(big[ch]->p, src + (size_t)ch * words, words * 8, cudaMemcpyHostToDevice, c.copy_stream));
        CNHE_CUDA(cudaEventRecord(c.ev_copy, c.copy_stream));
        CNHE_CUDA(cudaStreamWaitEvent(c.stream, c.ev_copy, 0));
    }
    for (int i = 0; i < n; i++) {
        cnhe_vec *o = new_vec(c, dim, scale, format, true, blocks);
        for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch); o->buf[ch] = big[ch]; o->off[ch] = (size_t)i * per; }
        out[i] = o;
    }
    API_END
}
extern "C" int cnhe_vecs_export_raw(cnhe_ctx *h, const cnhe_vec *const *vecs, int n, uint64_t *dst, size_t cap) {
    API_BEGIN(h)
    if (n < 1 || !dst) fail("bad arguments");
    const int blocks = vecs[0]->blocks;
    const size_t per = (size_t)blocks * c.ct_words();
    if (cap < (size_t)c.P * n * per) fail("destination too small");
    for (int i = 0; i < n; i++) {
        same_ctx(c, vecs[i]);
        if (!vecs[i]->enc || vecs[i]->blocks != blocks) fail("expecting encrypted vectors with equal block counts");
        for (int ch = 0; ch < c.P; ch++) {
            c.set_channel(ch);
            CNHE_CUDA(cudaMemcpyAsync(dst + ((size_t)ch * n + i) * per, vecs[i]->ptr(ch), per * 8, cudaMemcpyDeviceToHost, c.stream));
        }
    }
    c.sync();
    API_END
}
extern "C" int cnhe_vecs_export_raw_async(cnhe_ctx *h, const cnhe_vec *const *vecs, int n, uint64_t *dst, size_t cap, int *ticket) {
    API_BEGIN(h)
    if (n < 1 || !dst || !ticket) fail("bad arguments");
    const int blocks = vecs[0]->blocks;
    const size_t per = (size_t)blocks * c.ct_words();
    if (cap < (size_t)c.P * n * per) fail("destination too small");
    for (int i = 0; i < n; i++) {
        same_ctx(c, vecs[i]);
        if (!vecs[i]->enc || vecs[i]->blocks != blocks) fail("expecting encrypted vectors with equal block counts");
    }
    if (c.ev_export.empty()) {
        c.ev_export.resize((size_t)8 * c.P);
        for (cudaEvent_t &e : c.ev_export) CNHE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const int t = c.export_next;
    c.export_next = (c.export_next + 1) & 7;
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        for (int i = 0; i < n; i++)
            CNHE_CUDA(cudaMemcpyAsync(dst + ((size_t)ch * n + i) * per, vecs[i]->ptr(ch), per * 8, cudaMemcpyDeviceToHost, c.stream));
        CNHE_CUDA(cudaEventRecord(c.ev_export[(size_t)t * c.P + ch], c.stream));
    }
    *ticket = t;
    API_END
}
extern "C" int cnhe_export_wait(cnhe_ctx *h, int ticket) {
    if (!h) return set_err(CNHE_ERR_INVALID, "null context");
    Context &c = *h->c;
    try {
        if (ticket < 0 || ticket > 7 || c.ev_export.empty()) fail("bad ticket");
        // no context lock: waiting must not block another thread that is queueing the next batch
        CNHE_CUDA(cudaSetDevice(c.device));
        for (int ch = 0; ch < c.P; ch++) CNHE_CUDA(cudaEventSynchronize(c.ev_export[(size_t)ticket * c.P + ch]));
    }
    catch (const Error &e) { return set_err(e.code, e.what()); }
    catch (const std::exception &e) { return set_err(CNHE_ERR_INVALID, e.what()); }
    return CNHE_OK;
}
extern "C" int cnhe_vec_device_ptr(const cnhe_vec *v, int channel, uint64_t *dptr, size_t *words) {
    if (!v || channel < 0 || channel >= v->ctx->P) return set_err(CNHE_ERR_INVALID, "bad arguments");
    *dptr = (uint64_t)v->ptr(channel);
    if (words) *words = (size_t)v->blocks * v->unit();
    return CNHE_OK;
}
extern "C" int cnhe_noise_budget(cnhe_ctx *h, const cnhe_vec *v, int channel, int block, int *bits) {
    API_BEGIN(h)
    same_ctx(c, v);
    if (!v->enc || channel < 0 || channel >= c.P || block < 0 || block >= v->blocks) fail("bad arguments");
    *bits = op_noise_budget(c, channel, v->block(channel, block));
    API_END
}

static double log2_centred(u64 v, u64 t) { // log2 |v| of a residue mod t read as a centred integer
    const double d = v > t / 2 ? (double)(t - v) : (double)v;
    return d > 0 ? std::log2(d) : 0;
}
// evaluator.Add/Sub/AddMany launches that also feed the operation counters / noise trace (OperationsCount, CryptoTracker)
static void do_add(Context &c, int ch, const u64 *a, const u64 *b, u64 *out, size_t words, int sub) {
    c.check(launch_ct_add(a, b, out, words, c.k, c.logN, c.d_bc, sub, c.stream), sub ? "ct_sub" : "ct_add");
    c.note(sub ? Context::OP_SUB : Context::OP_ADD, ch, (int)(words / c.ct_words()), out, a, b);
}
static void do_add_many(Context &c, int ch, const std::vector<const u64 *> &terms, u64 *out) {
    const int n_in = (int)terms.size();
    c.check(launch_ct_add_many(upload_ptrs(c, terms), n_in, out, c.ct_words(), c.k, c.logN, c.d_bc, c.stream), "ct_add_many");
    c.op_count[Context::OP_ADD_MANY_ITEMS] += (uint64_t)n_in;
    double aux = 0; // tracing: log2 of the root-sum-square of the items' noise peaks 2^-(budget+1), i.e. the model's prediction input
    if (c.trace_noise) {
        double ss = 0;
        bool all = true;
        for (const u64 *t : terms) { const int b = c.known_budget(t); if (b < 0) { all = false; break; } ss += std::exp2(-2.0 * (b + 1)); }
        aux = all && ss > 0 ? 0.5 * std::log2(ss) : 0;
    }
    c.note(Context::OP_ADD_MANY, ch, 1, out, terms.empty() ? nullptr : terms[0], nullptr, aux);
    if (c.trace_noise && !c.trace.empty()) c.trace.back().n = n_in;
}

// ---------------------------------------------------------------------------------------------------- IVector operations
static void check_pair(const cnhe_vec *a, const cnhe_vec *b) {
    if (a->dim != b->dim) fail("Dimensions do not match");
    if (a->format != b->format) fail("Format mismatch");
}
// Add / Subtract (AtomicSealBfvVector.cs:983-1024, 1238-1271; wrapper EncryptedSealBfvVector.cs:271-282, 457-471)
static cnhe_vec *addsub(Context &c, const cnhe_vec *a, const cnhe_vec *b, bool sub) {
    if (!sub && a->scale == 0) return alias_of(b);
    if (b->scale == 0) return alias_of(a);
    if (a->scale != b->scale) fail("Scales do not match.");
    check_pair(a, b);
    if (!a->enc && !b->enc) fail("adding two plaintexts is not supported");
    if (sub && !a->enc) fail("the first argument for subtraction must be encrypted");
    const cnhe_vec *e = a->enc ? a : b, *p = a->enc ? b : a;
    if (e->blocks != p->blocks) fail("Dimensions do not match");
    cnhe_vec *o = new_vec(c, a->dim, a->scale, a->format, true, e->blocks);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        if (p->enc) {
            do_add(c, ch, a->ptr(ch), b->ptr(ch), o->ptr(ch), (size_t)e->blocks * c.ct_words(), sub);
        } else {
            const bool dense = p->format == CNHE_DENSE;
            c.check(launch_ct_add_plain(e->ptr(ch), o->ptr(ch), e->blocks, 2, p->ptr(ch), dense ? c.N : 1, dense ? (int)c.N : 1, c.k, c.logN, c.d_bc,
                                        c.ch[ch].pc, sub, c.stream),
                    "ct_add_plain");
            c.note(sub ? Context::OP_SUB_PLAIN : Context::OP_ADD_PLAIN, ch, e->blocks, o->ptr(ch), e->ptr(ch));
        }
    }
    return o;
}
extern "C" int cnhe_vec_add(cnhe_ctx *h, const cnhe_vec *a, const cnhe_vec *b, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a); same_ctx(c, b);
    *out = addsub(c, a, b, false);
    API_END
}
extern "C" int cnhe_vec_sub(cnhe_ctx *h, const cnhe_vec *a, const cnhe_vec *b, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a); same_ctx(c, b);
    *out = addsub(c, a, b, true);
    API_END
}

static std::vector<const u64 *> block_ptrs(const cnhe_vec *v, int ch, int repeat_first = 0) {
    std::vector<const u64 *> p;
    if (repeat_first) p.assign(repeat_first, v->block(ch, 0));
    else for (int b = 0; b < v->blocks; b++) p.push_back(v->block(ch, b));
    return p;
}
// PointwiseMultiplySparseDimOne (AtomicSealBfvVector.cs:774-810): `s` is the sparse dimension-one operand
static cnhe_vec *mul_sparse_dim_one(Context &c, const cnhe_vec *self, const cnhe_vec *s) {
    cnhe_vec *o = new_vec(c, self->dim, self->scale * s->scale, self->format, true, self->blocks);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        if (self->enc && s->enc) {
            op_multiply_relin(c, ch, block_ptrs(s, ch, self->blocks), block_ptrs(self, ch), o->ptr(ch));
        } else if (self->enc) { // constant plaintext times every block
            std::vector<u64> sc(self->blocks, s->scalars[ch][0]);
            if (sc[0] == 0) fail("plain cannot be zero (the result would be a transparent ciphertext)");
            u64 *d = c.ws_alloc(sc.size());
            CNHE_CUDA(cudaMemcpyAsync(d, sc.data(), sc.size() * 8, cudaMemcpyHostToDevice, c.stream));
            c.check(launch_ct_scale(self->ptr(ch), o->ptr(ch), self->blocks, 2, d, c.k, c.logN, c.d_bc, c.ch[ch].pc, c.stream), "ct_scale");
            c.note(Context::OP_MULTIPLY_SCALAR, ch, self->blocks, o->ptr(ch), self->ptr(ch), nullptr, log2_centred(sc[0], c.t[ch]));
            c.sync();
        } else { // plain blocks times the single ciphertext of s
            if (self->format != CNHE_DENSE) fail("unsupported plain format");
            u64 *rep = c.ws_alloc((size_t)self->blocks * c.ct_words());
            for (int b = 0; b < self->blocks; b++)
                { CNHE_CUDA(cudaMemcpyAsync(rep + (size_t)b * c.ct_words(), s->block(ch, 0), c.ct_words() * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(rep + (size_t)b * c.ct_words(), s->block(ch, 0)); }
            op_multiply_plain_dense(c, ch, rep, self->blocks, self->ptr(ch), true, o->ptr(ch));
        }
    }
    return guard.release();
}
// PointwiseMultiply (AtomicSealBfvVector.cs:813-860)
static cnhe_vec *pointwise_multiply(Context &c, const cnhe_vec *a, const cnhe_vec *b) {
    if (!a->enc && !b->enc) fail("multiplying two plaintexts is not implemented");
    if (a->dim == 1 && a->format == CNHE_SPARSE) return mul_sparse_dim_one(c, b, a);
    if (b->dim == 1 && b->format == CNHE_SPARSE) return mul_sparse_dim_one(c, a, b);
    check_pair(a, b);
    if (a->blocks != b->blocks) fail("Dimensions do not match");
    cnhe_vec *o = new_vec(c, a->dim, a->scale * b->scale, a->format, true, a->blocks);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    const cnhe_vec *e = a->enc ? a : b, *p = a->enc ? b : a;
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        if (a->enc && b->enc) {
            op_multiply_relin(c, ch, block_ptrs(b, ch), block_ptrs(a, ch), o->ptr(ch)); // evaluator.Multiply(ev.encData[i], encData[i])
        } else if (p->format == CNHE_DENSE) {
            op_multiply_plain_dense(c, ch, e->ptr(ch), e->blocks, p->ptr(ch), true, o->ptr(ch));
        } else {
            for (u64 s : p->scalars[ch])
                if (s == 0) fail("plain cannot be zero (the result would be a transparent ciphertext)");
            c.check(launch_ct_scale(e->ptr(ch), o->ptr(ch), e->blocks, 2, p->ptr(ch), c.k, c.logN, c.d_bc, c.ch[ch].pc, c.stream), "ct_scale");
            c.note(Context::OP_MULTIPLY_SCALAR, ch, e->blocks, o->ptr(ch), e->ptr(ch), nullptr, log2_centred(p->scalars[ch][0], c.t[ch]));
        }
    }
    return guard.release();
}
extern "C" int cnhe_vec_pointwise_multiply(cnhe_ctx *h, const cnhe_vec *a, const cnhe_vec *b, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a); same_ctx(c, b);
    *out = pointwise_multiply(c, a, b);
    API_END
}

// SumAllSlots (AtomicSealBfvVector.cs:888-955)
static cnhe_vec *sum_all_slots(Context &c, const cnhe_vec *a, uint64_t length, int force_column) {
    if (a->format != CNHE_DENSE) fail("Expecting dense vector format");
    if (length != CNHE_ALL_SLOTS && force_column >= 0) fail("forcing output in a column works only when doing complete sum");
    if (!a->enc) fail("SumAllSlots can be applied to encrypted data only");
    if (length == 0) fail("Can't sum over less then one element");
    if (length == 1) return alias_of(a);
    const size_t N = c.N, ctw = c.ct_words();
    uint64_t len = length;
    cnhe_vec *o = new_vec(c, a->dim, a->scale, CNHE_DENSE, true, 1);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        len = length;
        u64 *sum = o->ptr(ch);
        if (a->blocks > 1) { // AddMany over the blocks
            std::vector<const u64 *> ptrs = block_ptrs(a, ch);
            do_add_many(c, ch, ptrs, sum);
        } else {
            CNHE_CUDA(cudaMemcpyAsync(sum, a->ptr(ch), ctw * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(sum, a->ptr(ch));
        }
        u64 *tmp = c.ws_alloc(ctw);
        if (len >= N / 2) {
            if (!op_rotate_add(c, ch, sum, 1, 0, true, sum)) { // x += rotate(x) in one pass when the step has its own key
                op_rotate_columns(c, ch, sum, 1, tmp);
                do_add(c, ch, sum, tmp, sum, ctw, 0);
            }
            len = N / 2;
        }
        for (uint64_t steps = 1; steps < len; steps *= 2) { // RotateRowsAndAdd(sum, steps): RotateRows(c, -steps)
            if (op_rotate_add(c, ch, sum, 1, -(int)steps, false, sum)) continue;
            op_rotate_rows(c, ch, sum, 1, -(int)steps, tmp);
            do_add(c, ch, sum, tmp, sum, ctw, 0);
        }
        if (force_column >= 0) { // one-hot mask (":936-945")
            if ((size_t)force_column >= N) fail("column out of range");
            std::vector<u64> onehot(N, 0);
            onehot[force_column] = 1;
            u64 *dv = c.ws_alloc(N), *pl = c.ws_alloc(N);
            CNHE_CUDA(cudaMemcpyAsync(dv, onehot.data(), N * 8, cudaMemcpyHostToDevice, c.stream));
            op_encode(c, ch, dv, 1, (int)N, pl);
            op_multiply_plain_dense(c, ch, sum, 1, pl, false, sum);
            c.sync();
            len = 1;
        }
    }
    o->dim = (len >= N / 2) ? 1 : a->dim;
    o->format = (len >= N) ? CNHE_SPARSE : CNHE_DENSE;
    return guard.release();
}
extern "C" int cnhe_vec_sum_all_slots(cnhe_ctx *h, const cnhe_vec *a, uint64_t length, int force_column, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a);
    *out = sum_all_slots(c, a, length, force_column);
    API_END
}
extern "C" int cnhe_vec_dot_product(cnhe_ctx *h, const cnhe_vec *a, const cnhe_vec *b, uint64_t length, int force_column, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a); same_ctx(c, b);
    std::unique_ptr<cnhe_vec> mul(pointwise_multiply(c, a, b)); // AtomicSealBfvVector.cs:964-977
    *out = sum_all_slots(c, mul.get(), length, force_column);
    API_END
}
// Rotate (AtomicSealBfvVector.cs:1414-1430): first block only
extern "C" int cnhe_vec_rotate(cnhe_ctx *h, const cnhe_vec *a, int amount, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a);
    if (!a->enc) fail("Rotate operates only on encrypted data");
    if (a->format == CNHE_SPARSE) fail("Rotate operates only on dense vectors");
    cnhe_vec *o = new_vec(c, a->dim, a->scale, CNHE_DENSE, true, 1);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        op_rotate_rows(c, ch, a->ptr(ch), 1, amount, o->ptr(ch));
    }
    *out = guard.release();
    API_END
}
// Duplicate (AtomicSealBfvVector.cs:1370-1408)
extern "C" int cnhe_vec_duplicate(cnhe_ctx *h, const cnhe_vec *a, uint64_t count, cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a);
    uint64_t shift = 1;
    while (shift < a->dim) shift *= 2;
    if (!a->enc) fail("Duplicate operates only on encrypted data");
    if (a->format == CNHE_SPARSE) fail("Duplicate operates only on dense vectors");
    const size_t N = c.N, ctw = c.ct_words();
    if (shift * count > N) fail("Packed vector must fit in a single ciphertext");
    cnhe_vec *o = new_vec(c, count * shift, a->scale, CNHE_DENSE, true, 1);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        u64 *res = o->ptr(ch), *rotator = c.ws_alloc(ctw);
        CNHE_CUDA(cudaMemcpyAsync(res, a->ptr(ch), ctw * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(res, a->ptr(ch));
        const u64 *rot_src = a->ptr(ch);
        bool column_rotated = false;
        // the count-1 rotations of the original (or of its column-rotated copy) are independent: queue them, execute them together
        // (hops with the same Galois element share a key-switch wave), then add them in the reference's order (":1385-1402")
        std::vector<RotateJob> jobs;
        std::vector<u64 *> pieces;
        for (uint64_t i = 1; i < count; i++) {
            long long target = (long long)(i * shift);
            if (target * 2 >= (long long)N) {
                if (!column_rotated) {
                    column_rotated = true;
                    op_rotate_columns(c, ch, a->ptr(ch), 1, rotator);
                    rot_src = rotator;
                }
                target -= (long long)N / 2;
            }
            u64 *tmp = c.ws_alloc(ctw);
            jobs.push_back({rot_src, -(int)target, tmp}); // RotateRowsAndAdd(rotator, target, ...)
            pieces.push_back(tmp);
        }
        op_rotate_rows_multi(c, ch, jobs);
        for (u64 *tmp : pieces) do_add(c, ch, res, tmp, res, ctw, 0);
    }
    *out = guard.release();
    API_END
}
// Permute (AtomicSealBfvVector.cs:1436-1475)
extern "C" int cnhe_vec_permute(cnhe_ctx *h, const cnhe_vec *a, const cnhe_vec *const *selections, const int *shifts, int n, uint64_t output_dim,
                                cnhe_vec **out) {
    API_BEGIN(h)
    same_ctx(c, a);
    if (a->format != CNHE_DENSE) fail("Permute works only on dense vectors");
    if (!a->enc) fail("can permute only encrypted vectors");
    if (a->blocks > 1) fail("can permute only a single block");
    int first = -1;
    for (int i = 0; i < n; i++) {
        if (!selections[i]) continue;
        same_ctx(c, selections[i]);
        if (first < 0) first = i;
        if (selections[i]->dim != a->dim) fail("dimension of selection vector does not match dimension of data vector");
        if (selections[i]->scale != selections[first]->scale) fail("scales of all selection vectors should be the same");
        if (selections[i]->enc) fail("encrypted size must be 2 (rotating the size-3 product of an encrypted selection is rejected by SEAL)");
        if (selections[i]->format != CNHE_DENSE) fail("selection vectors must be dense");
    }
    if (first < 0) fail("permuting with no selected values is illigal");
    const size_t ctw = c.ct_words();
    cnhe_vec *o = new_vec(c, output_dim, a->scale * selections[first]->scale, CNHE_DENSE, true, 1);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        u64 *t = c.ws_alloc(ctw), *r = c.ws_alloc(ctw);
        bool have = false;
        for (int i = 0; i < n; i++) {
            if (!selections[i]) continue;
            op_multiply_plain_dense(c, ch, a->ptr(ch), 1, selections[i]->ptr(ch), false, t);
            op_rotate_rows(c, ch, t, 1, shifts[i], r);
            if (!have) { CNHE_CUDA(cudaMemcpyAsync(o->ptr(ch), r, ctw * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(o->ptr(ch), r); }
            else do_add(c, ch, o->ptr(ch), r, o->ptr(ch), ctw, 0);
            have = true;
        }
    }
    *out = guard.release();
    API_END
}
// GenerateSparseOfArray (AtomicSealBfvVector.cs:1347-1359)
extern "C" int cnhe_vecs_generate_sparse_of_array(cnhe_ctx *h, const cnhe_vec *const *vecs, int n, cnhe_vec **out) {
    API_BEGIN(h)
    if (n < 1) fail("empty array");
    for (int i = 0; i < n; i++) { same_ctx(c, vecs[i]); if (!vecs[i]->enc) fail("expecting encrypted vectors"); }
    cnhe_vec *o = new_vec(c, (uint64_t)n, vecs[0]->scale, CNHE_SPARSE, true, n);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        for (int i = 0; i < n; i++)
            { CNHE_CUDA(cudaMemcpyAsync(o->block(ch, i), vecs[i]->block(ch, 0), c.ct_words() * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(o->block(ch, i), vecs[i]->block(ch, 0)); }
    }
    *out = o;
    API_END
}

// Inteleave (AtomicSealBfvVector.cs:600-722): place vector k at slot offset shift*k
static void interleave_channel(Context &c, int ch, const std::vector<const cnhe_vec *> &vecs, int shift, int out_blocks, u64 *out) {
    const int block_size = (int)c.N, half = block_size / 2;
    const size_t ctw = c.ct_words();
    const int abs_shift = shift < 0 ? -shift : shift;
    if (shift < 0 && out_blocks > 1) fail("Negative shifts with multiple output blocks are not implemented yet");
    if (abs_shift > half && out_blocks > 1) fail("Shifts of more than half block size with multiple output blocks are not implemented yet");
    if ((long long)abs_shift * (long long)vecs.size() > (long long)block_size * out_blocks) fail("not enough room for interleaving");
    std::vector<std::vector<const u64 *>> lower(out_blocks), upper(out_blocks);
    auto ones_plain = [&](int count) { // BatchEncoder.Encode of `count` ones
        std::vector<u64> v(block_size, 0);
        for (int i = 0; i < count; i++) v[i] = 1;
        u64 *dv = c.ws_alloc(block_size), *pl = c.ws_alloc(block_size);
        CNHE_CUDA(cudaMemcpyAsync(dv, v.data(), (size_t)block_size * 8, cudaMemcpyHostToDevice, c.stream));
        op_encode(c, ch, dv, 1, block_size, pl);
        c.sync();
        return pl;
    };
    // phase 1: every vector's rotation (RotateRowsInplace of its own offset, ":625-660") -- independent of each other, so they are
    // queued and executed together: hops with the same Galois element share one key-switch wave (same per-ciphertext operations
    // and order as one rotate_rows call per vector, hence the same ciphertexts)
    struct Placed { int kind, start_block, end_block, in_block; u64 *v; }; // kind 0 lower, 1 upper, 2 straddles into the next block, 3 straddles lower/upper
    std::vector<Placed> placed(vecs.size());
    std::vector<RotateJob> jobs;
    for (size_t kk = 0; kk < vecs.size(); kk++) {
        long long this_shift = (long long)shift * (long long)kk;
        if (this_shift < 0) this_shift = half + this_shift;
        const int in_block = (int)(this_shift % block_size);
        const int start_block = (int)(this_shift / block_size), end_block = (int)((this_shift + abs_shift) / block_size);
        u64 *v = c.ws_alloc(ctw);
        const u64 *src = vecs[kk]->block(ch, 0);
        Placed &pl = placed[kk];
        pl.start_block = start_block; pl.end_block = end_block; pl.in_block = in_block; pl.v = v;
        if (in_block == 0) {
            CNHE_CUDA(cudaMemcpyAsync(v, src, ctw * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(v, src);
            pl.kind = 0;
        } else if (in_block + abs_shift < half) {
            jobs.push_back({src, -(int)this_shift, v});
            pl.kind = 0;
        } else if (in_block >= half) {
            jobs.push_back({src, -(in_block - half), v});
            pl.kind = start_block == end_block ? 1 : 2;
        } else {
            jobs.push_back({src, -in_block, v});
            pl.kind = 3;
        }
    }
    op_rotate_rows_multi(c, ch, jobs);
    // phase 2: split the vectors that straddle a half / block boundary with a one-hot-prefix mask (":640-672") and file every piece
    for (size_t kk = 0; kk < vecs.size(); kk++) {
        const Placed &pl = placed[kk];
        u64 *v = pl.v;
        const int start_block = pl.start_block, end_block = pl.end_block, in_block = pl.in_block;
        if (pl.kind == 0) {
            lower[start_block].push_back(v);
        } else if (pl.kind == 1) {
            upper[start_block].push_back(v);
        } else if (pl.kind == 2) { // straddles the upper half of this block and the lower half of the next
            const int upper_part = (in_block + abs_shift) - block_size;
            u64 *v2 = c.ws_alloc(ctw);
            CNHE_CUDA(cudaMemcpyAsync(v2, v, ctw * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(v2, v);
            op_multiply_plain_dense(c, ch, v, 1, ones_plain(upper_part), false, v);
            do_add(c, ch, v2, v, v2, ctw, 1);
            upper[start_block].push_back(v2);
            if (end_block >= out_blocks) fail("not enough room for interleaving");
            lower[end_block].push_back(v);
        } else { // straddles lower and upper half of the same block
            const int upper_part = (in_block + abs_shift) - half;
            if (upper_part > 0) {
                u64 *v2 = c.ws_alloc(ctw);
                CNHE_CUDA(cudaMemcpyAsync(v2, v, ctw * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(v2, v);
                op_multiply_plain_dense(c, ch, v, 1, ones_plain(upper_part), false, v);
                do_add(c, ch, v2, v, v2, ctw, 1);
                upper[start_block].push_back(v);
                lower[start_block].push_back(v2);
            } else {
                lower[start_block].push_back(v);
            }
        }
    }
    for (int i = 0; i < out_blocks; i++) {
        u64 *res = out + (size_t)i * ctw;
        if (lower[i].empty()) fail("an output block received no vector");
        do_add_many(c, ch, lower[i], res);
        if (!upper[i].empty()) {
            u64 *t = c.ws_alloc(ctw);
            do_add_many(c, ch, upper[i], t);
            op_rotate_columns(c, ch, t, 1, t);
            do_add(c, ch, res, t, res, ctw, 0);
        }
    }
}
static cnhe_vec *interleave(Context &c, const cnhe_vec *const *vecs, int n, int shift) { // AtomicSealBfvVector.cs:729-750
    if (n < 1) fail("empty array");
    std::vector<const cnhe_vec *> vv(vecs, vecs + n);
    for (auto v : vv) { same_ctx(c, v); if (!v->enc) fail("expecting encrypted vectors"); }
    if (vv[0]->format != CNHE_DENSE) fail("Expecting dense vector");
    int out_blocks = 1;
    if (shift > 0) out_blocks = (int)std::ceil((double)(vv[0]->dim * (uint64_t)n) / (double)c.N);
    cnhe_vec *o = new_vec(c, vv[0]->dim, vv[0]->scale, CNHE_DENSE, true, out_blocks);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        interleave_channel(c, ch, vv, shift, out_blocks, o->ptr(ch));
    }
    return guard.release();
}
extern "C" int cnhe_vecs_interleave(cnhe_ctx *h, const cnhe_vec *const *vecs, int n, int shift, cnhe_vec **out) {
    API_BEGIN(h)
    *out = interleave(c, vecs, n, shift);
    API_END
}
extern "C" int cnhe_vecs_stack(cnhe_ctx *h, const cnhe_vec *const *vecs, int n, cnhe_vec **out) { // AtomicSealBfvVector.cs:756-761
    API_BEGIN(h)
    if (n < 1) fail("empty array");
    cnhe_vec *o = interleave(c, vecs, n, (int)vecs[0]->dim);
    o->dim = vecs[0]->dim * (uint64_t)n;
    *out = o;
    API_END
}

// ---------------------------------------------------------------------------------------------------- matrix x vector, layers
struct RowHash {
    size_t operator()(const std::vector<int> &r) const {
        size_t hsh = 1469598103934665603ULL;
        for (int v : r) hsh = (hsh ^ (size_t)(unsigned)v) * 1099511628211ULL;
        return hsh;
    }
};
// ---- plan of a scalar-MAC layer for the tcgen05 kernel (mac_umma.cu): bundles of consecutive distinct gather rows
struct UmmaPlan {
    bool ok = false;
    std::vector<UmBundle> bundles;
    std::vector<int> chunk_rows;
    std::vector<unsigned char> wpack;
    std::vector<int> out_order;  // output index per (bundle, row)
    std::vector<int> extra_taps; // input index of every scratch-slab row (taps whose weights need the W2 part)
    int total_chunks = 0;
    // device copies (owned by the cache entry)
    UmBundle *d_bundles = nullptr;
    int *d_rows = nullptr;
    unsigned char *d_wpack = nullptr;
    ~UmmaPlan() {
        if (d_bundles) cudaFree(d_bundles);
        if (d_rows) cudaFree(d_rows);
        if (d_wpack) cudaFree(d_wpack);
    }
};
// bundles of `g` consecutive distinct rows; returns false when a bundle exceeds 128 outputs or a weight exceeds +-254
static bool umma_try(UmmaPlan &pl, int g, const std::vector<int> &grows, const std::vector<std::vector<int>> &row_outs, const std::vector<double> &wdh, int K) {
    const int R = (int)row_outs.size();
    std::unordered_map<std::string, int> seen;
    for (int start = 0; start < R; start += g) {
        const int end = std::min(R, start + g);
        std::vector<int> outs, out_row;
        int tmin = INT_MAX, tmax = -1;
        for (int r = start; r < end; r++) {
            for (int m : row_outs[r]) { outs.push_back(m); out_row.push_back(r); }
            for (int kk = 0; kk < K; kk++) {
                const int t = grows[(size_t)r * K + kk];
                if (t >= 0) { tmin = std::min(tmin, t); tmax = std::max(tmax, t); }
            }
        }
        if (outs.empty() || outs.size() > 128 || tmax < 0) return false;
        const int cols_main = ((tmax - tmin + 1) + 31) / 32 * 32;
        std::vector<int> wfull(outs.size() * (size_t)cols_main, 0);
        for (size_t i = 0; i < outs.size(); i++)
            for (int kk = 0; kk < K; kk++) {
                const int t = grows[(size_t)out_row[i] * K + kk];
                if (t >= 0) wfull[i * cols_main + (t - tmin)] += (int)wdh[(size_t)outs[i] * K + kk];
            }
        std::vector<int> extras; // columns that hold a weight beyond one signed byte
        for (int col = 0; col < cols_main; col++) {
            bool big = false;
            for (size_t i = 0; i < outs.size(); i++) {
                const int w = wfull[i * cols_main + col];
                if (w > 254 || w < -254) return false;
                big = big || w > 127 || w < -127;
            }
            if (big) extras.push_back(col);
        }
        const int cols = cols_main + ((int)extras.size() + 31) / 32 * 32;
        std::vector<signed char> w8(outs.size() * (size_t)cols, 0);
        for (size_t i = 0; i < outs.size(); i++) {
            for (int col = 0; col < cols_main; col++) w8[i * cols + col] = (signed char)std::max(-127, std::min(127, wfull[i * cols_main + col]));
            for (size_t j = 0; j < extras.size(); j++) {
                const int w = wfull[i * cols_main + extras[j]];
                w8[i * cols + cols_main + j] = (signed char)(w - std::max(-127, std::min(127, w)));
            }
        }
        std::string blob((size_t)cols / 32 * 4096, '\0');
        mac_umma_pack(w8.data(), (int)outs.size(), cols, reinterpret_cast<unsigned char *>(&blob[0]));
        UmBundle b;
        auto it = seen.find(blob);
        if (it == seen.end()) { // interior rows of a convolution share one matrix: the window slides with the bundle
            b.a_off = (int)pl.wpack.size();
            seen.emplace(blob, b.a_off);
            pl.wpack.insert(pl.wpack.end(), blob.begin(), blob.end());
        } else
            b.a_off = it->second;
        b.chunk0 = (int)pl.chunk_rows.size();
        b.n_chunks = cols / 32;
        b.n_out = (int)outs.size();
        b.out0 = (int)pl.out_order.size();
        for (int cch = 0; cch < cols_main / 32; cch++) pl.chunk_rows.push_back(tmin + 32 * cch);
        for (int cch = 0; cch < (cols - cols_main) / 32; cch++) pl.chunk_rows.push_back((1 << 30) | ((int)pl.extra_taps.size() + 32 * cch));
        for (int col : extras) pl.extra_taps.push_back(tmin + col);
        pl.out_order.insert(pl.out_order.end(), outs.begin(), outs.end());
        pl.bundles.push_back(b);
        pl.total_chunks += b.n_chunks;
    }
    return true;
}
// the search itself (host only): every bundle size g = 1, 2, ... rows; the plan with the fewest chunks per tile that fits in shared memory
static std::shared_ptr<UmmaPlan> umma_search(const std::vector<int> &grows, const std::vector<std::vector<int>> &row_outs, const std::vector<double> &wdh,
                                             int M, int K, int limbs) {
    std::shared_ptr<UmmaPlan> best;
    const int R = (int)row_outs.size();
    for (int g = 1; g <= R; g++) {
        auto pl = std::make_shared<UmmaPlan>();
        if (!umma_try(*pl, g, grows, row_outs, wdh, K)) {
            if (g > 1 && (size_t)g * row_outs[0].size() > 128) break; // larger groups only grow
            continue;
        }
        if (!mac_umma_fits((int)pl->wpack.size(), pl->total_chunks, M, (int)pl->bundles.size(), limbs)) continue;
        if (!best || pl->total_chunks < best->total_chunks) best = pl;
        if (g > 64) break;
    }
    return best;
}
// Planner probe for the CPU test suite (not part of include/cnhe.h, needs no GPU): gather[M][K] (-1 = padded tap), signed integer weights
// w[M][K]; out = {bundles, chunks per tile, weight bytes after de-duplication, extra (W2) taps}; returns 0 when no plan fits.
extern "C" int cnhe_debug_mac_plan(const int32_t *gather, const double *w, int M, int K, int limbs, int *out) {
    std::unordered_map<std::vector<int>, int, RowHash> index;
    std::vector<int> grows;
    std::vector<std::vector<int>> row_outs;
    for (int m = 0; m < M; m++) {
        std::vector<int> row(gather + (size_t)m * K, gather + (size_t)(m + 1) * K);
        auto it = index.find(row);
        if (it == index.end()) {
            index.emplace(row, (int)row_outs.size());
            grows.insert(grows.end(), row.begin(), row.end());
            row_outs.push_back({m});
        } else
            row_outs[it->second].push_back(m);
    }
    std::vector<double> wdh(w, w + (size_t)M * K);
    std::shared_ptr<UmmaPlan> pl = umma_search(grows, row_outs, wdh, M, K, limbs);
    if (!pl) return 0;
    out[0] = (int)pl->bundles.size(); out[1] = pl->total_chunks; out[2] = (int)pl->wpack.size(); out[3] = (int)pl->extra_taps.size();
    return 1;
}
// fewest chunks per tile over the bundle sizes that fit in shared memory; cached per layer (keyed by a hash of its gather table and weights)
static std::shared_ptr<UmmaPlan> umma_plan(Context &c, int ch, const std::vector<int> &grows, const std::vector<std::vector<int>> &row_outs,
                                           const std::vector<double> &wdh, int M, int K, int limbs) {
    u64 key = 0xcbf29ce484222325ULL ^ (u64)ch * 0x9e3779b97f4a7c15ULL ^ ((u64)M << 32) ^ (u64)K;
    auto mix = [&](const void *p, size_t bytes) {
        const u64 *w = reinterpret_cast<const u64 *>(p);
        for (size_t i = 0; i < bytes / 8; i++) { key ^= w[i]; key *= 0x100000001b3ULL; key ^= key >> 29; }
    };
    mix(wdh.data(), wdh.size() * 8);
    {
        std::vector<int> shape(grows); // the distinct gather rows, then which row every output reads
        shape.resize(grows.size() + (size_t)M + 1, -2);
        for (size_t r = 0; r < row_outs.size(); r++)
            for (int m : row_outs[r]) shape[grows.size() + (size_t)m] = (int)r;
        mix(shape.data(), shape.size() / 2 * 8);
    }
    auto hit = c.umma_plans.find(key);
    if (hit != c.umma_plans.end()) return std::static_pointer_cast<UmmaPlan>(hit->second);
    std::shared_ptr<UmmaPlan> best = umma_search(grows, row_outs, wdh, M, K, limbs);
    if (!best) best = std::make_shared<UmmaPlan>();
    else {
        best->ok = true;
        CNHE_CUDA(cudaMalloc((void **)&best->d_bundles, best->bundles.size() * sizeof(UmBundle)));
        CNHE_CUDA(cudaMalloc((void **)&best->d_rows, best->chunk_rows.size() * sizeof(int)));
        CNHE_CUDA(cudaMalloc((void **)&best->d_wpack, best->wpack.size()));
        CNHE_CUDA(cudaMemcpy(best->d_bundles, best->bundles.data(), best->bundles.size() * sizeof(UmBundle), cudaMemcpyHostToDevice));
        CNHE_CUDA(cudaMemcpy(best->d_rows, best->chunk_rows.data(), best->chunk_rows.size() * sizeof(int), cudaMemcpyHostToDevice));
        CNHE_CUDA(cudaMemcpy(best->d_wpack, best->wpack.data(), best->wpack.size(), cudaMemcpyHostToDevice));
    }
    if (c.umma_plans.size() > 64) c.umma_plans.clear();
    c.umma_plans[key] = best;
    return best;
}
// Shared body of DenseMatrixBySparseVectorMultiply (ciphertext columns x plain constants) and of the fused PoolLayer.
// in[n_in] encrypted dense vectors (same block count), weights[m] plain sparse of dim K, bias[m] plain dense or null.
static void mac_layer(Context &c, const cnhe_vec *const *in, int n_in, const int32_t *gather, const cnhe_vec *const *weights, const cnhe_vec *const *bias,
                      int M, int K, cnhe_vec **out) {
    if (n_in < 1 || M < 1 || K < 1) fail("bad layer shape");
    const int bl = in[0]->blocks;
    for (int i = 0; i < n_in; i++) {
        same_ctx(c, in[i]);
        if (!in[i]->enc || in[i]->format != CNHE_DENSE) fail("layer inputs must be encrypted dense vectors");
        if (in[i]->blocks != bl || in[i]->dim != in[0]->dim || in[i]->scale != in[0]->scale) fail("all layer inputs must share dimension and scale");
    }
    bool const_bias = bias != nullptr;
    for (int m = 0; m < M; m++) {
        same_ctx(c, weights[m]);
        if (weights[m]->enc || weights[m]->format != CNHE_SPARSE) fail("expecting a sparse vector");
        if (weights[m]->dim != (uint64_t)K) fail("dimensions do not match");
        if (weights[m]->scale != weights[0]->scale) fail("weight scales differ");
        if (bias) {
            if (!bias[m]) fail("null bias");
            same_ctx(c, bias[m]);
            if (bias[m]->enc || bias[m]->format != CNHE_DENSE || bias[m]->dim != in[0]->dim) fail("bias must be a plain dense vector of the input dimension");
            if (bias[m]->scale != in[0]->scale * weights[0]->scale) fail("Scales do not match.");
            const_bias = const_bias && bias[m]->is_const;
        }
    }
    // 128-bit accumulator bound: K products of (q_l - 1)^2
    int maxbits = 0;
    for (u64 q : c.q) maxbits = std::max(maxbits, hm::bit_length(q));
    if (2 * maxbits + hm::bit_length((u64)K) > 127) fail("layer too wide for the 128-bit accumulator with these coefficient moduli");
    // tiles: outputs that share a gather row, 8 at a time
    std::vector<int> grows;
    std::vector<MacTile> tiles;
    std::vector<std::vector<int>> row_outs; // outputs of every distinct gather row, in the order of `grows`
    {
        std::unordered_map<std::vector<int>, std::vector<int>, RowHash> groups;
        std::vector<std::vector<int>> order;
        for (int m = 0; m < M; m++) {
            std::vector<int> row(K);
            bool any = false;
            for (int kk = 0; kk < K; kk++) {
                row[kk] = gather ? gather[(size_t)m * K + kk] : kk;
                if (row[kk] >= n_in) fail("gather index out of range");
            }
            // The reference skips zero plaintexts per plaintext modulus (IsZero, AtomicSealBfvVector.cs:468) and sums what is left, so an
            // output whose taps are all = 0 modulo one of the t_c has an empty sum in that channel.  (Deviation, documented in DESIGN.md:
            // padded taps are skipped here, while the reference multiplies a fresh encryption of zero by their weight.)
            for (int ch = 0; ch < c.P; ch++) {
                bool any_ch = false;
                for (int kk = 0; kk < K; kk++) any_ch = any_ch || (row[kk] >= 0 && weights[m]->scalars[ch][kk] != 0);
                any = any || any_ch;
                if (!any_ch) fail("an output has no non-zero tap modulo one of the plaintext primes (the reference would sum an empty list)");
            }
            if (!any) fail("an output has no non-zero tap (the reference would sum an empty list)");
            auto it = groups.find(row);
            if (it == groups.end()) { order.push_back(row); groups[row] = {m}; }
            else it->second.push_back(m);
        }
        for (auto &row : order) {
            const int row_index = (int)(grows.size() / K);
            grows.insert(grows.end(), row.begin(), row.end());
            const std::vector<int> &ms = groups[row];
            row_outs.push_back(ms);
            for (size_t s = 0; s < ms.size(); s += 8) {
                MacTile t;
                memset(&t, 0, sizeof(t));
                t.gather_row = row_index;
                t.n_out = (int)std::min<size_t>(8, ms.size() - s);
                for (int j = 0; j < t.n_out; j++) t.out_index[j] = ms[s + j];
                tiles.push_back(t);
            }
        }
    }
    const int order_rows = (int)(grows.size() / K);
    const double out_scale = in[0]->scale * weights[0]->scale;
    std::vector<BufRef> big(c.P);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        big[ch] = c.alloc((size_t)M * bl * c.ct_words());
    }
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        // every temporary lives on this channel's stream (stream-ordered frees must not overtake another channel's kernels)
        int *d_gather = (int *)c.ws_alloc((grows.size() + 1) / 2 + 1);
        c.h2d(d_gather, grows.data(), grows.size() * sizeof(int));
        MacTile *d_tiles = (MacTile *)c.ws_alloc((tiles.size() * sizeof(MacTile) + 7) / 8);
        c.h2d(d_tiles, tiles.data(), tiles.size() * sizeof(MacTile));
        std::vector<const u64 *> wp(M);
        for (int m = 0; m < M; m++) wp[m] = weights[m]->ptr(ch);
        const u64 *const *d_w = upload_ptrs(c, wp);
        // signed weights as doubles for the FP64 accumulate path, when they are small enough for it to be exact
        const u64 t = c.t[ch], thr = (t + 1) >> 1;
        std::vector<double> wdh((size_t)M * K);
        double wmax = 0;
        for (int m = 0; m < M; m++)
            for (int kk = 0; kk < K; kk++) {
                const u64 w = weights[m]->scalars[ch][kk];
                const double d = w >= thr ? -(double)(t - w) : (double)w;
                wdh[(size_t)m * K + kk] = d;
                wmax = std::max(wmax, std::fabs(d));
            }
        const bool fp_mac = maxbits <= 50 && wmax < 131072.0 && (double)K * wmax * 67108864.0 < 4503599627370496.0 && !getenv("CNHE_MAC_INT");
        // dense layer (one gather row shared by every output, 8-bit weights): exact integer GEMM on the tensor cores (mac_imma.cu)
        const int limbs = (maxbits + 7) / 8;
        const bool imma = order_rows == 1 && wmax <= 254.0 && K >= 32 && M >= 8 && limbs >= 5 && limbs <= 7 && (double)K * 254.0 * 255.0 < 2147483648.0 &&
                          !getenv("CNHE_MAC_NO_IMMA") && !getenv("CNHE_MAC_INT");
        // ... or, for any layer whose inputs are evenly spaced rows of one slab (the previous layer's output, an imported batch) and whose
        // weights stay within +-254: tcgen05 (mac_umma.cu), dense and convolution alike
        std::shared_ptr<UmmaPlan> plan;
        long long tap_stride = 0;
        {
            // (M >= 8: LoLa's per-map products come one output at a time -- a 128-row MMA per tile would be 99 % padding and the kernel's
            // per-tile latency more than the whole scalar-MAC launch)
            bool slab = bl == 1 && n_in >= 2 && M >= 8 && limbs >= 5 && limbs <= 7 && wmax <= 254.0 && !getenv("CNHE_MAC_NO_UMMA") && !getenv("CNHE_MAC_NO_IMMA") &&
                        !getenv("CNHE_MAC_INT");
            if (slab) {
                tap_stride = in[1]->block(ch, 0) - in[0]->block(ch, 0);
                slab = tap_stride >= (long long)c.ct_words() && tap_stride % 2 == 0;
                for (int i = 0; i < n_in && slab; i++) slab = in[i]->block(ch, 0) == in[0]->block(ch, 0) + (long long)i * tap_stride;
            }
            if (slab) plan = umma_plan(c, ch, grows, row_outs, wdh, M, K, limbs);
        }
        const bool umma = plan && plan->ok && (double)plan->total_chunks * 32.0 * 127.0 * 255.0 < 2147483648.0;
        const void *d_wfrag = nullptr, *d_wfrag2 = nullptr;
        if (!umma && imma) {
            const int mtiles = (M + 15) / 16, chunks = (K + 31) / 32;
            const size_t fwords = (size_t)mtiles * chunks * 32 * 4;
            std::vector<uint32_t> frag(2 * fwords, 0); // W1 = clamp(W, +-127) then the residual W2 = W - W1
            bool any2 = false;
            auto wq = [&](int m, int kk, int part) -> uint32_t { // signed weight of output m, tap kk (0 outside the matrix / on padded taps)
                if (m >= M || kk >= K || grows[kk] < 0) return 0;
                const int w = (int)wdh[(size_t)m * K + kk], w1 = std::max(-127, std::min(127, w));
                const int v = part ? w - w1 : w1;
                if (part && v) any2 = true;
                return (uint32_t)(uint8_t)(int8_t)v;
            };
            for (int part = 0; part < 2; part++)
                for (int mt = 0; mt < mtiles; mt++)
                    for (int cch = 0; cch < chunks; cch++)
                        for (int lane = 0; lane < 32; lane++) {
                            const int g = lane >> 2, tig = lane & 3;
                            uint32_t *dst = &frag[part * fwords + (((size_t)mt * chunks + cch) * 32 + lane) * 4];
                            for (int r = 0; r < 4; r++) { // r: 0 (row g, k 0..15) 1 (row g+8) 2 (row g, k 16..31) 3 (row g+8, k 16..31)
                                const int row = mt * 16 + g + (r & 1) * 8, k0 = cch * 32 + tig * 4 + (r >> 1) * 16;
                                dst[r] = wq(row, k0, part) | (wq(row, k0 + 1, part) << 8) | (wq(row, k0 + 2, part) << 16) | (wq(row, k0 + 3, part) << 24);
                            }
                        }
            u64 *buf = c.ws_alloc((frag.size() * 4 + 7) / 8);
            c.h2d(buf, frag.data(), (any2 ? 2 : 1) * fwords * 4);
            d_wfrag = buf;
            if (any2) d_wfrag2 = reinterpret_cast<const uint32_t *>(buf) + fwords;
        }
        const double *d_wd = nullptr;
        if (fp_mac) {
            u64 *buf = c.ws_alloc(wdh.size());
            c.h2d(buf, wdh.data(), wdh.size() * 8);
            d_wd = reinterpret_cast<const double *>(buf);
        }
        const u64 *d_bias = nullptr;
        if (bias && const_bias) {
            std::vector<u64> bv(M);
            for (int m = 0; m < M; m++) bv[m] = bias[m]->const_val[ch];
            u64 *db = c.ws_alloc(M);
            c.h2d(db, bv.data(), (size_t)M * 8);
            d_bias = db;
        }
        for (int b = 0; b < bl; b++) {
            std::vector<const u64 *> ip(n_in);
            for (int i = 0; i < n_in; i++) ip[i] = in[i]->block(ch, b);
            std::vector<u64 *> op(M);
            for (int m = 0; m < M; m++) op[m] = big[ch]->p + ((size_t)m * bl + b) * c.ct_words();
            double used = 0;
            for (auto &t : tiles) { int kk = 0; for (int j = 0; j < K; j++) kk += grows[(size_t)t.gather_row * K + j] >= 0; used += kk + t.n_out; }
            c.prof_begin(4, used * 8.0 * c.ct_words());
            if (umma) {
                UmmaLaunch a;
                memset(&a, 0, sizeof(a));
                a.slab = in[0]->block(ch, 0);
                a.slab_stride_words = (size_t)tap_stride;
                a.slab_rows = (size_t)n_in;
                if (!plan->extra_taps.empty()) { // the taps that need the W2 part, side by side
                    u64 *scratch = c.ws_alloc(plan->extra_taps.size() * c.ct_words());
                    for (size_t j = 0; j < plan->extra_taps.size(); j++)
                        CNHE_CUDA(cudaMemcpyAsync(scratch + j * c.ct_words(), ip[plan->extra_taps[j]], c.ct_words() * 8, cudaMemcpyDeviceToDevice, c.stream));
                    a.scratch = scratch;
                    a.scratch_rows = plan->extra_taps.size();
                }
                std::vector<u64 *> opo(M);
                for (int i = 0; i < M; i++) opo[i] = op[plan->out_order[i]];
                const u64 *d_bias_o = nullptr;
                if (d_bias) {
                    std::vector<u64> bo(M);
                    for (int i = 0; i < M; i++) bo[i] = bias[plan->out_order[i]]->const_val[ch];
                    u64 *db = c.ws_alloc(M);
                    c.h2d(db, bo.data(), (size_t)M * 8);
                    d_bias_o = db;
                }
                a.bundles = plan->d_bundles; a.n_bundles = (int)plan->bundles.size();
                a.chunk_rows = plan->d_rows; a.total_chunks = plan->total_chunks;
                a.wpack = plan->d_wpack; a.a_bytes = (int)plan->wpack.size();
                a.out_ptrs = upload_ptrs_mut(c, opo); a.bias = d_bias_o; a.n_out_total = M;
                a.limbs = limbs; a.k = c.k; a.logn = c.logN; a.bc = c.d_bc; a.pc = c.ch[ch].pc;
                c.check(launch_mac_umma(a, c.stream), "mac_umma");
            } else if (imma) {
                std::vector<const u64 *> ipg(K);
                for (int kk = 0; kk < K; kk++) ipg[kk] = ip[grows[kk] < 0 ? 0 : grows[kk]]; // padded taps carry weight 0
                c.check(launch_mac_dense_imma(upload_ptrs(c, ipg), d_wfrag, d_wfrag2, d_bias, K, M, limbs, upload_ptrs_mut(c, op), c.k, c.logN, c.d_bc, c.ch[ch].pc,
                                              c.stream),
                        "mac_dense_imma");
            } else if (fp_mac)
                c.check(launch_mac_layer_fp(upload_ptrs(c, ip), d_gather, d_tiles, (int)tiles.size(), d_wd, d_bias, K, upload_ptrs_mut(c, op), c.k, c.logN,
                                            c.d_bc, c.ch[ch].pc, c.stream),
                        "mac_layer_fp");
            else
                c.check(launch_mac_layer(upload_ptrs(c, ip), d_gather, d_tiles, (int)tiles.size(), d_w, d_bias, K, upload_ptrs_mut(c, op), c.k, c.logN,
                                         c.d_bc, c.ch[ch].pc, c.stream),
                        "mac_layer");
            c.prof_end();
        }
        { // what the reference issues for this layer (AtomicSealBfvVector.cs:466-475, 497-505): MultiplyPlain per non-zero tap, AddMany per output
            uint64_t taps = 0;
            for (int m = 0; m < M; m++)
                for (int kk = 0; kk < K; kk++) taps += (!gather || gather[(size_t)m * K + kk] >= 0) && weights[m]->scalars[ch][kk] != 0;
            c.op_count[Context::OP_MULTIPLY_SCALAR] += taps * bl;
            c.op_count[Context::OP_ADD_MANY_ITEMS] += taps * bl;
            if (bias) c.op_count[Context::OP_ADD_PLAIN] += (uint64_t)M * bl;
            double ss = 0; // output 0: root-sum-square of its centred weights (the noise gain of the scalar MAC under independent inputs)
            for (int kk = 0; kk < K; kk++)
                if (!gather || gather[kk] >= 0) ss += wdh[kk] * wdh[kk];
            c.note(Context::OP_ADD_MANY, ch, M * bl, big[ch]->p, in[gather ? std::max(gather[0], 0) : 0]->block(ch, 0), nullptr, ss > 0 ? 0.5 * std::log2(ss) : 0);
        }
        if (bias && !const_bias) // generic AddPlain per output
            for (int m = 0; m < M; m++) {
                u64 *o = big[ch]->p + (size_t)m * bl * c.ct_words();
                c.check(launch_ct_add_plain(o, o, bl, 2, bias[m]->ptr(ch), c.N, (int)c.N, c.k, c.logN, c.d_bc, c.ch[ch].pc, 0, c.stream), "ct_add_plain");
            }
    }
    for (int m = 0; m < M; m++) {
        cnhe_vec *o = new_vec(c, in[0]->dim, out_scale, CNHE_DENSE, true, bl);
        for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
            o->buf[ch] = big[ch];
            o->off[ch] = (size_t)m * bl * c.ct_words();
        }
        out[m] = o;
    }
}
extern "C" int cnhe_layer_conv_dense(cnhe_ctx *h, const cnhe_vec *const *in, int n_in, const int32_t *gather, const cnhe_vec *const *weights,
                                     const cnhe_vec *const *bias, int M, int K, cnhe_vec **out) {
    API_BEGIN(h)
    mac_layer(c, in, n_in, gather, weights, bias, M, K, out);
    API_END
}
// DenseMatrixBySparseVectorMultiply (AtomicSealBfvVector.cs:434-521)
extern "C" int cnhe_mat_mul_colmajor_sparse(cnhe_ctx *h, const cnhe_vec *const *cols, int K, const cnhe_vec *sparse, cnhe_vec **out) {
    API_BEGIN(h)
    if (K < 1) fail("empty matrix");
    same_ctx(c, sparse);
    for (int i = 0; i < K; i++) same_ctx(c, cols[i]);
    if ((uint64_t)K != sparse->dim) fail("dimensions do not match");
    if (sparse->format != CNHE_SPARSE) fail("expecting a sparse vector");
    if (!cols[0]->enc && !sparse->enc) fail("at least one parameter has to be encrypted");
    if (cols[0]->enc && !sparse->enc) {
        mac_layer(c, cols, K, nullptr, &sparse, nullptr, 1, K, out);
    } else {
        const int bl = cols[0]->blocks;
        const size_t ctw = c.ct_words();
        cnhe_vec *o = new_vec(c, cols[0]->dim, cols[0]->scale * sparse->scale, CNHE_DENSE, true, bl);
        std::unique_ptr<cnhe_vec> guard(o);
        alloc_channels(o);
        for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
            u64 *prod = c.ws_alloc((size_t)K * bl * ctw);
            if (cols[0]->enc) { // both encrypted: Multiply + Relinearize per (k, block)
                std::vector<const u64 *> a, b;
                for (int kk = 0; kk < K; kk++)
                    for (int i = 0; i < bl; i++) { a.push_back(cols[kk]->block(ch, i)); b.push_back(sparse->block(ch, kk)); }
                op_multiply_relin(c, ch, a, b, prod);
            } else { // plain columns x encrypted constants: MultiplyPlain(sparse.enc[k], denses[k].plain[i])
                for (int kk = 0; kk < K; kk++) {
                    u64 *rep = c.ws_alloc((size_t)bl * ctw);
                    for (int i = 0; i < bl; i++)
                        { CNHE_CUDA(cudaMemcpyAsync(rep + (size_t)i * ctw, sparse->block(ch, kk), ctw * 8, cudaMemcpyDeviceToDevice, c.stream)); c.note_copy(rep + (size_t)i * ctw, sparse->block(ch, kk)); }
                    op_multiply_plain_dense(c, ch, rep, bl, cols[kk]->ptr(ch), true, prod + (size_t)kk * bl * ctw);
                }
            }
            for (int i = 0; i < bl; i++) {
                std::vector<const u64 *> terms;
                for (int kk = 0; kk < K; kk++) terms.push_back(prod + ((size_t)kk * bl + i) * ctw);
                do_add_many(c, ch, terms, o->block(ch, i));
            }
            c.sync();
        }
        *out = guard.release();
    }
    API_END
}
// Batched SumAllSlots (AtomicSealBfvVector.cs:888-955) on n single-block ciphertexts in place: the same rotate-and-add ladder,
// one key-switch wave per step for all n.
static uint64_t sum_slots_batched(Context &c, int ch, u64 *cts, int n, uint64_t length) {
    const size_t N = c.N, words = (size_t)n * c.ct_words();
    uint64_t len = length;
    u64 *tmp = c.ws_alloc(words);
    // every step is x += rotate(x): fused into the rotation (the permutation kernel folds x into the key switch's base) when the step
    // has its own Galois key -- it does for the powers of two the ladder walks -- else rotate, then add
    if (len >= N / 2) {
        if (!op_rotate_add(c, ch, cts, n, 0, true, cts)) {
            op_rotate_columns(c, ch, cts, n, tmp);
            do_add(c, ch, cts, tmp, cts, words, 0);
        }
        len = N / 2;
    }
    for (uint64_t steps = 1; steps < len; steps *= 2) {
        if (op_rotate_add(c, ch, cts, n, -(int)steps, false, cts)) continue;
        op_rotate_rows(c, ch, cts, n, -(int)steps, tmp);
        do_add(c, ch, cts, tmp, cts, words, 0);
    }
    return len;
}
// RowMajor matrix x vector (EncryptedSealBfvMatrix.cs:79-120): per row DotProduct(row, v) = PointwiseMultiply + SumAllSlots, then
// GenerateSparseOfArray, or (ForceDenseFormat) a one-hot mask per row and the sum of all rows.  All rows go through each stage
// together instead of one DotProduct per row.
extern "C" int cnhe_mat_mul_rowmajor(cnhe_ctx *h, const cnhe_vec *const *rows, int n_rows, const cnhe_vec *v, int force_dense, cnhe_vec **out) {
    return cnhe_mat_mul_rowmajor_shard(h, rows, n_rows, v, force_dense, 0, n_rows, out);
}
// The same product for a contiguous SLICE of the matrix rows (rows[i] is global row first_row + i of a matrix with total_rows rows):
// the unit of the intra-inference multi-GPU split (CIFAR's 5488 dense rows over 4 GPUs).  ForceDense: the one-hot masks sit
// at the global columns, so the partial results of the ranks add up to the full product (the reference sums the masked rows too,
// EncryptedSealBfvMatrix.cs:92-116); otherwise the output holds this slice's sparse elements only.
extern "C" int cnhe_mat_mul_rowmajor_shard(cnhe_ctx *h, const cnhe_vec *const *rows, int n_rows, const cnhe_vec *v, int force_dense, int first_row,
                                           int total_rows, cnhe_vec **out) {
    API_BEGIN(h)
    if (n_rows < 1) fail("empty matrix");
    if (first_row < 0 || first_row + n_rows > total_rows) fail("row slice out of range");
    same_ctx(c, v);
    if (!v->enc) fail("at least one parameter has to be encrypted");
    if (v->format != CNHE_DENSE) fail("Expecting dense vector format");
    if (v->blocks != 1) fail("row-major multiplication expects a single-block vector");
    for (int r = 0; r < n_rows; r++) {
        same_ctx(c, rows[r]);
        if (rows[r]->enc) fail("encrypted rows are not supported by the batched row-major product");
        if (rows[r]->dim != v->dim) fail("Dimensions do not match");
        if (rows[r]->format != v->format) fail("Format mismatch");
        if (rows[r]->scale != rows[0]->scale) fail("row scales differ");
    }
    const size_t N = c.N, ctw = c.ct_words();
    if (force_dense && (size_t)total_rows > N) fail("column out of range");
    const int out_blocks = force_dense ? 1 : n_rows;
    cnhe_vec *o = new_vec(c, (uint64_t)(force_dense ? total_rows : n_rows), v->scale * rows[0]->scale, force_dense ? CNHE_DENSE : CNHE_SPARSE, true, out_blocks);
    std::unique_ptr<cnhe_vec> guard(o);
    alloc_channels(o);
    const int RC = 1024; // rows per wave
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        bool first = true;
        for (int r0 = 0; r0 < n_rows; r0 += RC) {
            WsScope scope(c);
            const int m = std::min(RC, n_rows - r0);
            u64 *plains = c.ws_alloc((size_t)m * N);
            for (int i = 0; i < m; i++)
                CNHE_CUDA(cudaMemcpyAsync(plains + (size_t)i * N, rows[r0 + i]->ptr(ch), N * 8, cudaMemcpyDeviceToDevice, c.stream));
            u64 *prod = force_dense ? c.ws_alloc((size_t)m * ctw) : o->block(ch, r0);
            op_multiply_plain_dense_bcast(c, ch, v->ptr(ch), plains, m, prod);
            sum_slots_batched(c, ch, prod, m, CNHE_ALL_SLOTS);
            if (force_dense) {
                // one-hot masks for columns r0..r0+m (EncryptedSealBfvMatrix.cs:96, AtomicSealBfvVector.cs:936-945)
                u64 *masks = c.ws_alloc((size_t)m * N);
                op_encode_onehot(c, ch, m, first_row + r0, masks); // built on the device (was a 134 MB pageable upload per 1024-row wave)
                op_multiply_plain_dense(c, ch, prod, m, masks, true, prod);
                std::vector<const u64 *> terms;
                if (!first) terms.push_back(o->ptr(ch));
                for (int i = 0; i < m; i++) terms.push_back(prod + (size_t)i * ctw);
                do_add_many(c, ch, terms, o->ptr(ch));
                c.sync();
            }
            first = false;
        }
    }
    *out = guard.release();
    API_END
}
// SquareActivation over a whole matrix: every column PointwiseMultiply'd with itself in one wave per channel
extern "C" int cnhe_layer_square(cnhe_ctx *h, const cnhe_vec *const *in, int n, cnhe_vec **out) {
    API_BEGIN(h)
    if (n < 1) fail("empty layer");
    std::vector<int> first(n + 1, 0);
    for (int i = 0; i < n; i++) {
        same_ctx(c, in[i]);
        if (!in[i]->enc) fail("multiplying two plaintexts is not implemented");
        first[i + 1] = first[i] + in[i]->blocks;
    }
    const int total = first[n];
    std::vector<BufRef> big(c.P);
    for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
        big[ch] = c.alloc((size_t)total * c.ct_words());
        std::vector<const u64 *> ptrs;
        for (int i = 0; i < n; i++)
            for (int b = 0; b < in[i]->blocks; b++) ptrs.push_back(in[i]->block(ch, b));
        op_multiply_relin(c, ch, ptrs, ptrs, big[ch]->p);
    }
    for (int i = 0; i < n; i++) {
        cnhe_vec *o = new_vec(c, in[i]->dim, in[i]->scale * in[i]->scale, in[i]->format, true, in[i]->blocks);
        for (int ch = 0; ch < c.P; ch++) {
        c.set_channel(ch);
            o->buf[ch] = big[ch];
            o->off[ch] = (size_t)first[i] * c.ct_words();
        }
        out[i] = o;
    }
    API_END
}
