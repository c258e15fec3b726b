// Host runtime of libcnhe: context (device tables, keys, workspace) and the ciphertext-array operations the vector
// layer (vec.cu) is built from.  Mirrors AtomicSealBfvEncryptedEnvironment ("HE Wrapper/AtomicSealBfvVector.cs:19-206")
// plus the SEAL objects it owns (SEALContext, KeyGenerator, Evaluator, Encryptor, Decryptor, BatchEncoder).
#pragma once
#include <cstdint>
#include <algorithm>
#include <map>
#include <unordered_map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "kernels.h"

namespace cnhe {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
void cuda_check(cudaError_t e, const char *what);
#define CNHE_CUDA(x) ::cnhe::cuda_check((x), #x)

// Reference-counted device allocation (stream-ordered pool).
struct DevBuf {
    u64 *p = nullptr;
    size_t words = 0;
    cudaStream_t stream = nullptr; // release stream: the channel the buffer belongs to (also the allocation stream unless given)
    struct Context *owner = nullptr; // set: a large block goes back to the context's per-stream recycle list instead of the driver pool
    DevBuf(size_t w, cudaStream_t s);
    DevBuf(size_t w, cudaStream_t s, struct Context *owner);
    int upload_slot = -1; // >= 0: the block is one of the context's persistent upload slots (returned, not freed)
    ~DevBuf();
    DevBuf(const DevBuf &) = delete;
};
typedef std::shared_ptr<DevBuf> BufRef;

struct Channel { // one plaintext modulus (one AtomicSealBfvEncryptedEnvironment)
    u64 t = 0;
    PlainConst pc;
    int mod_id = 0; // NTT table id of t
    bool have_sk = false, have_pk = false, have_rlk = false;
    BufRef sk, pk, rlk;
    std::map<u64, BufRef> glk;
    RngKey rng;    // secure (ChaCha20 keyed from the OS) unless a deterministic test seed was requested explicitly
    u64 nonce = 1; // running encryption counter (32 bits enter the stream id; a secure channel re-keys before it wraps)
    FloorConstF floor_f; // folded fast_floor constants for this t (valid when the context's fp_elementwise is set)
};

struct Context {
    int device = 0;
    uint32_t N = 0;
    int logN = 0, k = 0, kb = 0, P = 0, dbc_relin = 0, dbc_galois = 0; // kb: primes in the BEHZ base Bsk
    std::vector<u64> q, bsk, t;
    BehzConst h_bc;
    BehzConst *d_bc = nullptr;
    BehzConstF h_bf;
    BehzConstF *d_bf = nullptr;
    bool fp_elementwise = false; // every q_i, Bsk prime small enough for the FP64 element-wise kernels
    bool lazy = false;           // ... and for lazy-double intermediates between the kernels of a multiply / key switch
    std::vector<NttTab> h_tabs;
    NttTab *d_tabs = nullptr;
    u64 *d_table_mem = nullptr;
    std::vector<u32> h_index_map;
    u32 *d_index_map = nullptr;
    DigitMap dm_relin, dm_galois;
    std::vector<u64> galois_elts;
    std::vector<Channel> ch;
    // one CUDA stream per plaintext-modulus channel (the reference runs one Task per prime, EncryptedSealBfvVector.cs:225-236):
    // channels are independent until decryption, so their kernels and host<->device copies overlap.  `stream` is the stream of the
    // channel currently being issued (set_channel).
    std::vector<cudaStream_t> streams;
    cudaStream_t stream = nullptr;
    bool multi_stream = true;
    void set_channel(int ch) { stream = streams[multi_stream ? ch : 0]; }
    void join_streams(); // stream 0 waits for the tail of every other stream
    void fork_streams(); // every other stream waits for the tail of stream 0
    // bulk ciphertext uploads (cnhe_vecs_import_raw) run on their own stream, fenced by events against the owning channel's stream:
    // the upload of the next batch overlaps the kernels of the current one (double buffering across API calls)
    cudaStream_t copy_stream = nullptr;
    // persistent device blocks for uploaded ciphertext batches: a slot is handed out again once the event recorded at its release (on
    // the consuming channel's stream) allows it -- no driver allocation in the steady state, and with three slots in rotation the
    // upload of batch i+1 never waits for batch i's kernels
    struct UploadSlot { u64 *p; size_t words; cudaEvent_t released; bool busy; uint64_t stamp; };
    std::vector<UploadSlot> upload_slots;
    uint64_t upload_stamp = 0;
    BufRef alloc_upload(size_t words, cudaStream_t release_stream); // the copy stream is made to wait for the slot's last release
    void release_upload(int slot, cudaStream_t s);
    cudaEvent_t ev_copy = nullptr;
    std::vector<cudaEvent_t> ev_export; // ring of 8 tickets x P channel events (cnhe_vecs_export_raw_async)
    int export_next = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_join = nullptr;
    std::recursive_mutex mu;
    std::vector<BufRef> temps; // workspace temporaries of the operation in flight (guarded by mu)
    // ---- operation counters (the reference's OperationsCount, "HE Wrapper/AtomicSealBfvVector.cs:211-294") and the optional
    // per-operation noise-budget trace (CryptoTracker.TestBudget, "HE Wrapper/CryptoTracker.cs:41-52")
    enum OpKind { OP_ENCRYPT, OP_DECRYPT, OP_MULTIPLY, OP_RELINEARIZE, OP_MULTIPLY_PLAIN, OP_MULTIPLY_SCALAR, OP_ADD, OP_ADD_PLAIN, OP_SUB,
                  OP_SUB_PLAIN, OP_ROTATE_ROWS_HOP, OP_ROTATE_COLUMNS, OP_ADD_MANY, OP_ADD_MANY_ITEMS, OP_COUNT };
    uint64_t op_count[OP_COUNT] = {0};
    bool trace_noise = false;
    struct TraceRec { int kind, channel, n, budget, in0, in1, aux_milli, reserved; };
    std::vector<TraceRec> trace;
    std::map<const u64 *, int> budget_of; // tracing only: last measured budget of the ciphertext at a device address
    // count `n` operations of `kind`; with tracing on, also record the invariant noise budget of the first output ciphertext next to the
    // budgets its first input ciphertexts had (so that every operation can be checked against the analytic noise model on its own) and
    // an operation-specific `aux` value (log2 of the scalar / of the root-sum-square weight of a MAC output)
    void note(OpKind kind, int channel, int n, const u64 *first_out = nullptr, const u64 *in0 = nullptr, const u64 *in1 = nullptr, double aux = 0);
    void note_copy(const u64 *dst, const u64 *src) { // a device-to-device copy of a ciphertext keeps its budget
        if (!trace_noise) return;
        auto it = budget_of.find(src);
        if (it != budget_of.end()) budget_of[dst] = it->second; else budget_of.erase(dst);
    }
    int known_budget(const u64 *p) const { auto it = budget_of.find(p); return it == budget_of.end() ? -1 : it->second; }
    int chunk = 1024; // ciphertexts per kernel wave (upper bound: wave() also keeps a wave's scratch under ~8 GiB)
    int wave(size_t words_per_ct) const { // ciphertexts per wave for an operation needing `words_per_ct` scratch words per ciphertext
        const size_t fit = ((size_t)1 << 30) / (words_per_ct ? words_per_ct : 1); // 2^30 words = 8 GiB
        // with one stream per plaintext modulus the channels' kernels interleave on the GPU: 128-ciphertext waves keep that interleaving
        // fine grained (measured: 28.8 ms per pipelined batch against 35.2 ms with whole-layer waves); a single stream prefers one wave
        const size_t cap = multi_stream && streams.size() > 1 ? std::min(chunk, 128) : chunk;
        return (int)std::max<size_t>(16, std::min<size_t>(cap, fit));
    }
    uint64_t launches = 0;
    // optional per-kernel-family timing
    bool prof = false;
    struct ProfRec { int family; double bytes; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[6] = {0, 0, 0, 0, 0, 0}, prof_bytes[6] = {0, 0, 0, 0, 0, 0};
    uint64_t prof_n[6] = {0, 0, 0, 0, 0, 0};
    void prof_begin(int family, double bytes);
    void prof_end();
    void prof_flush();
    size_t ws_used = 0; // words handed out as temporaries since the last release (diagnostic)
    // host CRT data of the wrapper ("HE Wrapper/EncryptedSealBfvVector.cs:79-90")
    unsigned __int128 big_factor = 0;
    std::vector<unsigned __int128> crt_coeff;

    // pinned staging ring for small host->device uploads (pointer tables, weights, tiles): truly asynchronous copies
    unsigned char *stage_buf = nullptr;
    size_t stage_size = 0, stage_off = 0;
    // the ring is cut into STAGE_PARTS parts; leaving a part records one event per channel stream, entering it waits for the events of
    // its previous use (several parts ago: already complete in the steady state) -- no device-wide sync when the ring wraps
    static constexpr int STAGE_PARTS = 8;
    std::vector<cudaEvent_t> stage_ev; // [part][stream]
    std::vector<char> stage_ev_set;
    void h2d(void *dst, const void *src, size_t bytes); // async on `stream`; `src` may be freed on return

    ~Context();
    size_t ct_words() const { return (size_t)2 * k * N; }
    u64 *ws_alloc(size_t words); // temporary of the current operation: released (stream ordered / recycled) by WsScope or the next API call
    BufRef alloc(size_t words) { return std::make_shared<DevBuf>(words, stream, this); }
    // Large blocks (layer slabs, the 1 GB digit waves) are recycled per stream: a block released on stream S is handed to the next
    // request of a similar size on S without a driver call -- stream order makes that safe, and it removes cudaMallocAsync's slow path
    // (measured: 120 ms for 1 GB when the pool has no fitting free block) from the steady state.
    struct Recycled { u64 *p; size_t words; cudaStream_t stream; };
    std::vector<Recycled> recycle;
    size_t recycle_words = 0;
    bool recycle_on = true;
    u64 *take_recycled(size_t words, cudaStream_t s, size_t &got_words);
    bool give_recycled(u64 *p, size_t words, cudaStream_t s);
    void drop_recycled();
    void launched(int n = 1) { launches += n; }
    void check(cudaError_t e, const char *what) { cuda_check(e, what); launched(); if (trace_ms > 0) trace_gap(what); }
    std::unordered_map<u64, std::shared_ptr<void>> umma_plans; // tcgen05 layer plans (vec.cu), keyed by a hash of the layer's weights and gather table
    double trace_ms = 0; // CNHE_TRACE_SLOW: report host-side gaps between consecutive launches longer than this
    void trace_gap(const char *what);
    void sync();
};

void ws_release_all(Context &c); // drop every workspace temporary (call at the start of a public operation)
struct WsScope {                  // temporaries allocated inside the scope are released when it ends
    Context &c;
    size_t mark;
    explicit WsScope(Context &ctx);
    ~WsScope();
};

Context *context_create(const u64 *plain_primes, int P, uint32_t N, const u64 *coeff, int k, int dbc_relin, int dbc_galois, int device);
std::vector<u64> default_coeff_modulus(uint32_t N);

// ---- keys
void keys_generate(Context &c, u64 seed); // deterministic sampler: tests only
void keys_generate_secure(Context &c);    // fresh OS entropy per channel
void rng_from_os(RngKey &rk);
const char *op_kind_name(int kind);
BufRef &key_slot(Context &c, int channel, int what, u64 arg, size_t &words, bool create);

// ---- ciphertext-array operations (all asynchronous on c.stream; device pointers)
// upload a host array of device pointers into workspace memory
const u64 *const *upload_ptrs(Context &c, const std::vector<const u64 *> &ptrs);
u64 *const *upload_ptrs_mut(Context &c, const std::vector<u64 *> &ptrs);

void op_ntt(Context &c, const u64 *src, u64 *dst, int n_polys, int mod_base, int mod_count, bool inverse);
// out3[n][3][k][N] = a[i] * b[i]  (BEHZ).  a_ptrs/b_ptrs: host vectors of device ciphertext pointers.
void op_multiply(Context &c, int ch, const std::vector<const u64 *> &a, const std::vector<const u64 *> &b, u64 *out3);
void op_relinearize(Context &c, int ch, const u64 *in3, int n, u64 *out2);
void op_multiply_relin(Context &c, int ch, const std::vector<const u64 *> &a, const std::vector<const u64 *> &b, u64 *out2);
void op_key_switch(Context &c, const u64 *target, size_t target_stride, int n, const u64 *key, const DigitMap &dm, const u64 *base,
                   size_t base_stride, u64 *out);
void op_apply_galois(Context &c, int ch, const u64 *in, int n, u64 elt, u64 *out, bool add_back = false);
bool op_rotate_add(Context &c, int ch, const u64 *in, int n, int steps, bool columns, u64 *out); // out = in + rotate(in) in one pass, if possible
void op_rotate_rows(Context &c, int ch, const u64 *in, int n, int steps, u64 *out); // steps == 0 copies
void op_rotate_columns(Context &c, int ch, const u64 *in, int n, u64 *out);
// Many independent single-ciphertext row rotations with DIFFERENT step counts (Interleave / Stack / Duplicate rotate every vector by its
// own offset): each job walks the hop sequence rotate_rows would take for it (exact key or NAF hops), and hops with the same Galois
// element are batched across jobs into one key-switch wave.  Per ciphertext the operations and their order are exactly those of
// op_rotate_rows, so the outputs are bit-identical.
struct RotateJob { const u64 *src; int steps; u64 *dst; };
void op_rotate_rows_multi(Context &c, int ch, const std::vector<RotateJob> &jobs);
u64 galois_elt_from_step(const Context &c, int steps);
// dense plaintext (coefficient form mod t, [n or 1][N]) times ciphertexts [n][2kN]
void op_multiply_plain_dense(Context &c, int ch, const u64 *ct, int n, const u64 *plain, bool plain_per_ct, u64 *out);
void op_multiply_plain_dense_bcast(Context &c, int ch, const u64 *ct, const u64 *plains, int n, u64 *out);
// values [n][count] (mod t, device) -> plain [n][N] coefficient form
void op_encode(Context &c, int ch, const u64 *values, int n, int count, u64 *plain);
void op_decode(Context &c, int ch, const u64 *plain, int n, u64 *values);
// plain[i] = BatchEncoder.Encode(e_(first_col + i)) for i < n, built on the device (the one-hot masks of ForceOutputInColumn)
void op_encode_onehot(Context &c, int ch, int n, int first_col, u64 *plain);
// plain [n][plain_stride] (first `coeffs` coefficients used) -> ct [n][2kN]; nonces nonce0..nonce0+n-1
// reserve n consecutive encryption nonces of a channel (a secure channel re-keys from the OS before the 32-bit counter wraps)
u64 take_nonces(Context &c, int ch, u64 n);
void op_encrypt(Context &c, int ch, const u64 *plain, size_t plain_stride, int n, int coeffs, u64 nonce0, u64 *ct);
void op_decrypt(Context &c, int ch, const u64 *ct, int n, u64 *plain);
int op_noise_budget(Context &c, int ch, const u64 *ct);

} // namespace cnhe
