/*
 * cnhe.h -- C ABI of libcnhe.so, the B200-native BFV engine behind the CryptoNets plugin API.
 *
 * This is the drop-in boundary.  A C# `B200BfvFactory : IFactory` binds these entry points
 * with [DllImport("cnhe")] (stub in INTEGRATION.md); the Python mirror in cryptonets_b200/ binds them with ctypes.
 * One cnhe_vec is one reference `EncryptedSealBfvVector` ("HE Wrapper/EncryptedSealBfvVector.cs:150-573"): P
 * plaintext-modulus channels, each an `AtomicSealBfvEncryptedVector` ("HE Wrapper/AtomicSealBfvVector.cs:303-1476")
 * whose SEAL Ciphertext[] / Plaintext[] live in B200 HBM.  Every call is batched over blocks and channels;
 * nothing here computes on the CPU and the library fails at load/first call when no CUDA device is present.
 *
 * Conventions: every function returns 0 (CNHE_OK) or a negative error code; cnhe_last_error() gives the message of
 * the last failure on the calling thread (the reference throws System.Exception with the same wording where it has
 * one).  Nothing throws across the ABI.  Handles are opaque; vectors are immutable after creation except for the
 * metadata setters; destroying a vector that another vector/matrix aliases is safe (buffers are reference counted,
 * mirroring `CopyVectors:false` / `DataDisposedExternaly` in "HE Wrapper/EncryptedSealBfvMatrix.cs:32-58").
 * Calls may come from several host threads (reference: one IComputationEnvironment per thread,
 * "HE Wrapper/Utils.cs:46-88"); work is serialised onto the context's CUDA stream.
 *
 * Raw ciphertext layout (import/export): SEAL's in-memory layout, [poly][residue][coefficient] uint64, coefficient
 * (non-NTT) form, canonical residues.
 */
#ifndef CNHE_H
#define CNHE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CNHE_OK 0
#define CNHE_ERR_INVALID (-1)   /* bad argument / dimension / format / scale mismatch (reference: System.Exception) */
#define CNHE_ERR_CUDA (-2)      /* CUDA runtime failure, including "no device" */
#define CNHE_ERR_STATE (-3)     /* keys missing, etc. */
#define CNHE_ERR_UNSUPPORTED (-4)

#define CNHE_DENSE 0  /* EVectorFormat.dense  ("HE Wrapper/IVector.cs:15-18") */
#define CNHE_SPARSE 1 /* EVectorFormat.sparse */
#define CNHE_ALL_SLOTS 0x7fffffffu /* Int32.MaxValue length of SumAllSlots ("AtomicSealBfvVector.cs:873,880") */

typedef struct cnhe_ctx cnhe_ctx; /* == EncryptedSealBfvFactory + its reference EncryptedSealBfvEnvironment */
typedef struct cnhe_vec cnhe_vec; /* == EncryptedSealBfvVector */

const char *cnhe_last_error(void);
const char *cnhe_version(void);

/* ---- context & keys ------------------------------------------------------------------------------------------ */
/* new EncryptedSealBfvFactory(primes, n, DecompositionBitCount, GaloisDecompositionBitCount, SmallModulusCount)
 * ("HE Wrapper/IFactory.cs:255-260"); coefficient modulus = first small_modulus_count entries (<=0: all) of
 * DefaultParams.CoeffModulus128(N) ("AtomicSealBfvVector.cs:140-151").  Builds every device table; no keys yet. */
int cnhe_context_create(const uint64_t *plain_primes, int P, uint32_t N, int dbc_relin, int dbc_galois,
                        int small_modulus_count, int device, cnhe_ctx **out);
/* same with explicit coefficient moduli (AtomicSealBfvVector.cs:152-161 Parms(t, n, coefModulus)) */
int cnhe_context_create_custom(const uint64_t *plain_primes, int P, uint32_t N, const uint64_t *coeff_moduli, int k,
                               int dbc_relin, int dbc_galois, int device, cnhe_ctx **out);
int cnhe_context_destroy(cnhe_ctx *);
int cnhe_context_info(const cnhe_ctx *, uint32_t *N, int *k, int *P, int *relin_digits, int *galois_digits, int *galois_elts);
int cnhe_context_coeff_moduli(const cnhe_ctx *, uint64_t *out_k);
int cnhe_context_plain_moduli(const cnhe_ctx *, uint64_t *out_P);
/* the BEHZ auxiliary base Bsk (auxiliary primes, m_sk last); out may be NULL to query the count.  NTT modulus ids:
 * 0..k-1 coefficient primes, k..k+count-1 Bsk, k+count+c plaintext modulus c */
int cnhe_context_bsk_moduli(const cnhe_ctx *, uint64_t *out, int *count);
int cnhe_context_galois_elts(const cnhe_ctx *, uint64_t *out);
/* options: "behz_centered_mtilde" (0/1), "chunk" (ciphertexts per multiply/key-switch wave), "multi_stream" (1: one CUDA
 * stream per plaintext modulus, default; 0: everything on one stream; refused while imported batches are alive),
 * "trace_noise" (1: record the invariant noise budget after every evaluator-level operation, see cnhe_trace_read) */
int cnhe_context_set_option(cnhe_ctx *, const char *name, int64_t value);
int cnhe_context_sync(cnhe_ctx *);
/* interop with the caller's own GPU work (the NCCL all-gather of the score ciphertexts): the CUDA stream (cudaStream_t as an integer) of a
 * channel, and the fences of the per-channel streams -- join: stream 0 waits for every channel's tail; fork: every channel waits for
 * stream 0.  Work the caller enqueues on stream 0 between a join and a fork is ordered after everything queued before the join and
 * before everything queued after the fork. */
int cnhe_context_stream(cnhe_ctx *, int channel, uint64_t *stream);
int cnhe_context_join_streams(cnhe_ctx *);
int cnhe_context_fork_streams(cnhe_ctx *);
/* EncryptedSealBfvEnvironment.GenerateEncryptionKeys ("EncryptedSealBfvVector.cs:92-102") -> KeyGenerator, RelinKeys(dbc),
 * GaloisKeys(dbc) ("AtomicSealBfvVector.cs:62-74"), sampled on the device.
 * cnhe_keys_generate_secure: every channel draws a fresh 256-bit ChaCha20 key from the OS (getrandom) -- secret key, key masks and all
 * later encryption randomness come from it (what SEAL's std::random_device gives the reference).  This is the production call.
 * A context is born with such a key per channel, so encrypting under an imported public key is safe without any seeding call.
 * cnhe_keys_generate(seed): DETERMINISTIC, TESTS ONLY -- the counter-based splitmix64 sampler shared with the CPU oracle (channel c uses
 * seed + c), so keys and fresh ciphertexts are bit-comparable; its outputs are predictable from the public key. */
int cnhe_keys_generate_secure(cnhe_ctx *);
int cnhe_keys_generate(cnhe_ctx *, uint64_t seed);
/* what: 0 secret key [k][N] (NTT form), 1 public key [2][k][N], 2 relin keys [D][2][k][N], 3 Galois key of element
 * `arg` [D][2][k][N].  Stand-in for the SEAL key streams of SaveToStream/LoadFromStream ("AtomicSealBfvVector.cs:93-130"). */
int cnhe_keys_export(cnhe_ctx *, int channel, int what, uint64_t arg, uint64_t *dst, size_t cap_words);
int cnhe_keys_import(cnhe_ctx *, int channel, int what, uint64_t arg, const uint64_t *src, size_t words);
int cnhe_keys_set_seed(cnhe_ctx *, int channel, uint64_t seed); /* TESTS ONLY: later encryptions of that channel use the deterministic sampler */
/* OperationsCount ("HE Wrapper/AtomicSealBfvVector.cs:211-294"): evaluator-level operations issued since creation / the last reset, in
 * the order Encryption, Decryption, Multiplication, Relinarization, PlainMultiplication (dense plaintext), ScalarMultiplication (constant
 * plaintext: SEAL's monomial multiply_plain), Addition, PlainAddition, Subtraction, PlainSubtraction, Rotation (row-rotation hops),
 * ColumnRotation, AddMany, AddManyItemCount.  cnhe_op_name(i) names counter i. */
#define CNHE_OP_COUNT 14
int cnhe_op_counts(cnhe_ctx *, uint64_t *out, int cap, int reset);
const char *cnhe_op_name(int kind);
/* CryptoTracker.TestBudget ("HE Wrapper/CryptoTracker.cs:41-52") as a trace: with option "trace_noise" on, every evaluator-level
 * operation appends a record of eight int32: kind, channel, count, invariant noise budget of its first output ciphertext (-1 when not
 * measured), the budgets its first and second input ciphertexts had when they were last measured (-1 unknown), an operation-specific
 * auxiliary value in thousandths (log2 |scalar| of a constant multiply, log2 of the root-sum-square weight of a MAC output, log2 of the
 * root-sum-square input noise of an AddMany) and a reserved word.  out may be NULL to query the record count.  Switching the option off
 * forgets the per-ciphertext budgets. */
int cnhe_trace_read(cnhe_ctx *, int32_t *out, size_t cap_records, size_t *n_records, int clear);

/* ---- wire / on-disk formats.  The containers are the reference's; the SEAL 3.2 binary streams inside them are
 * restated from knowledge of SEAL 3.2.x and are UNPINNED against the real binary (see csrc/wire.cu for every layout).
 * cnhe_keys_save: EncryptedSealBfvEnvironment.Save ("HE Wrapper/EncryptedSealBfvVector.cs:104-134", IFactory.Save "IFactory.cs:484-495"):
 *   ZIP archive with one `environmentNNN` entry per plaintext modulus = AtomicSealBfvEncryptedEnvironment.SaveToStream
 *   ("AtomicSealBfvVector.cs:93-104").  dst may be NULL to query the size.
 * cnhe_context_load: the factory's file constructor (EncryptedSealBfvFactory(fileName), LoadFromStream "AtomicSealBfvVector.cs:106-131"):
 *   parameters (N, coefficient moduli, plaintext moduli, decomposition bit counts) and keys come from the archive; a context loaded
 *   from an archive without secret keys can encrypt and evaluate but not decrypt.
 * cnhe_vec_write / cnhe_vec_read: EncryptedSealBfvVector.Write / Read (":414-439", "AtomicSealBfvVector.cs:1273-1345"), the text form
 *   IFactory.LoadVector / LoadMatrix parse ("IFactory.cs:474-483"; a matrix is the reference's three header lines around its vectors). */
int cnhe_keys_save(cnhe_ctx *, int with_private_keys, uint8_t *dst, size_t cap, size_t *needed);
int cnhe_context_load(const uint8_t *archive, size_t len, int device, cnhe_ctx **out);
int cnhe_vec_write(cnhe_ctx *, const cnhe_vec *, char *dst, size_t cap, size_t *needed);
int cnhe_vec_read(cnhe_ctx *, const char *text, size_t len, cnhe_vec **out, size_t *consumed);

/* ---- vectors: creation, metadata, disposal ---------------------------------------------------------------------- */
/* IFactory.GetEncryptedVector / GetPlainVector ("HE Wrapper/IFactory.cs:311-328"): round(v*scale), CRT split over the
 * plain primes ("EncryptedSealBfvVector.cs:352-365"), BatchEncoder.Encode per N-slot block (dense) or one constant
 * polynomial per element (sparse) ("AtomicSealBfvVector.cs:1114-1142"), Encryptor.Encrypt (":1202-1216"). */
int cnhe_vec_encrypt(cnhe_ctx *, const double *v, uint64_t dim, double scale, int format, cnhe_vec **out);
int cnhe_vec_plain(cnhe_ctx *, const double *v, uint64_t dim, double scale, int format, cnhe_vec **out);
/* the BigInteger overloads IFactory.GetPlainVector / GetEncryptedVector(IEnumerable<BigInteger>, format) ("IFactory.cs:29,43";
 * "EncryptedSealBfvVector.cs:188-199"): residues [P][dim], already reduced modulo each plaintext prime by the caller (SplitBigNumbers) */
int cnhe_vec_from_residues(cnhe_ctx *, const uint64_t *residues, uint64_t dim, double scale, int format, int encrypt, cnhe_vec **out);
/* batched form of IFactory.GetEncryptedMatrix ("IFactory.cs:353-380"): n dense vectors of `dim` values, v row-major [n][dim] */
int cnhe_vecs_encrypt(cnhe_ctx *, const double *v, int n, uint64_t dim, double scale, cnhe_vec **out);
/* IVector.Decrypt ("EncryptedSealBfvVector.cs:332-337,381-395"; "AtomicSealBfvVector.cs:1030-1067") */
int cnhe_vec_decrypt(cnhe_ctx *, const cnhe_vec *, double *out, uint64_t cap);
int cnhe_vecs_decrypt(cnhe_ctx *, const cnhe_vec *const *vecs, int n, double *out /*[n][dim]*/, uint64_t dim);
/* IVector.DecryptFullPrecision ("EncryptedSealBfvVector.cs:343-348"): per-channel residues [P][dim]; the caller joins them with big
 * integers (JoinSplitNumbers ":397-411") */
int cnhe_vec_decrypt_residues(cnhe_ctx *, const cnhe_vec *, uint64_t *out, uint64_t cap_words);
int cnhe_vec_copy(cnhe_ctx *, const cnhe_vec *, cnhe_vec **out); /* IFactory.CopyVector */
int cnhe_vec_destroy(cnhe_vec *);
/* Dispose of n vectors in one call (the Dispose loop of EncryptedSealBfvMatrix, EncryptedSealBfvMatrix.cs Dispose); null entries are skipped. */
int cnhe_vecs_destroy(cnhe_vec *const *vecs, int n);
int cnhe_vec_meta(const cnhe_vec *, uint64_t *dim, double *scale, int *format, int *is_encrypted, int *blocks, uint64_t *block_size);
int cnhe_vec_register_scale(cnhe_vec *, double scale); /* IVector.RegisterScale */
int cnhe_vec_register_dim(cnhe_vec *, uint64_t dim);   /* AtomicSealBfvVector.cs:316-319 */
/* raw ciphertext access for parity tests: block `block` of channel `channel`, 2*k*N words */
int cnhe_vec_export_raw(cnhe_ctx *, const cnhe_vec *, int channel, int block, uint64_t *dst, size_t cap_words);
int cnhe_vec_import_raw(cnhe_ctx *, const uint64_t *src /*[P][blocks][2kN]*/, int blocks, uint64_t dim, double scale, int format,
                        cnhe_vec **out);
/* batched forms: n vectors of `blocks` ciphertexts each; host layout [P][n][blocks][2kN].  One copy per channel; the
 * host buffer may be pinned (cudaHostAlloc / torch pin_memory) for full PCIe rate. */
int cnhe_vecs_import_raw(cnhe_ctx *, const uint64_t *src, int n, int blocks, uint64_t dim, double scale, int format, cnhe_vec **out);
int cnhe_vecs_export_raw(cnhe_ctx *, const cnhe_vec *const *vecs, int n, uint64_t *dst, size_t cap_words);
/* Asynchronous form for a pipelined serving loop: the device-to-host copies are queued behind the kernels that produce the vectors
 * and the call returns at once with a ticket; cnhe_export_wait(ticket) blocks until exactly those copies have landed in `dst`
 * (pinned host memory), without waiting for work queued afterwards.  Up to 8 tickets may be outstanding. */
int cnhe_vecs_export_raw_async(cnhe_ctx *, const cnhe_vec *const *vecs, int n, uint64_t *dst, size_t cap_words, int *ticket);
int cnhe_export_wait(cnhe_ctx *, int ticket);
/* device pointer of a channel's ciphertext blocks (for NCCL gathers through torch; plumbing only) */
int cnhe_vec_device_ptr(const cnhe_vec *, int channel, uint64_t *dptr, size_t *words);
int cnhe_noise_budget(cnhe_ctx *, const cnhe_vec *, int channel, int block, int *bits); /* CryptoTracker.cs:41-52 */

/* ---- IVector operations ------------------------------------------------------------------------------------------ */
int cnhe_vec_add(cnhe_ctx *, const cnhe_vec *a, const cnhe_vec *b, cnhe_vec **out);          /* AtomicSealBfvVector.cs:983-1024 */
int cnhe_vec_sub(cnhe_ctx *, const cnhe_vec *a, const cnhe_vec *b, cnhe_vec **out);          /* :1238-1271 */
int cnhe_vec_pointwise_multiply(cnhe_ctx *, const cnhe_vec *a, const cnhe_vec *b, cnhe_vec **out); /* :813-860, :774-810 */
/* length == CNHE_ALL_SLOTS: full sum; force_column < 0: none */
int cnhe_vec_sum_all_slots(cnhe_ctx *, const cnhe_vec *a, uint64_t length, int force_column, cnhe_vec **out); /* :888-955 */
int cnhe_vec_dot_product(cnhe_ctx *, const cnhe_vec *a, const cnhe_vec *b, uint64_t length, int force_column, cnhe_vec **out); /* :964-977 */
int cnhe_vec_rotate(cnhe_ctx *, const cnhe_vec *a, int amount, cnhe_vec **out);              /* :1414-1430 */
int cnhe_vec_duplicate(cnhe_ctx *, const cnhe_vec *a, uint64_t count, cnhe_vec **out);       /* :1370-1408 */
int cnhe_vec_permute(cnhe_ctx *, const cnhe_vec *a, const cnhe_vec *const *selections, const int *shifts, int n,
                     uint64_t output_dim, cnhe_vec **out);                                   /* :1436-1475 */
int cnhe_vecs_interleave(cnhe_ctx *, const cnhe_vec *const *vecs, int n, int shift, cnhe_vec **out); /* :600-750 */
int cnhe_vecs_stack(cnhe_ctx *, const cnhe_vec *const *vecs, int n, cnhe_vec **out);         /* :756-761 */
int cnhe_vecs_generate_sparse_of_array(cnhe_ctx *, const cnhe_vec *const *vecs, int n, cnhe_vec **out); /* :1347-1359 */

/* ---- IMatrix.Mul and the fused layer entry points ------------------------------------------------------------------ */
/* ColumnMajor matrix x sparse vector ("EncryptedSealBfvMatrix.cs:70-78" -> "AtomicSealBfvVector.cs:434-521") */
int cnhe_mat_mul_colmajor_sparse(cnhe_ctx *, const cnhe_vec *const *cols, int K, const cnhe_vec *sparse, cnhe_vec **out);
/* RowMajor plain matrix x encrypted dense vector ("EncryptedSealBfvMatrix.cs:79-120" -> DotProduct per row = MultiplyPlain +
 * SumAllSlots, "AtomicSealBfvVector.cs:964-977,888-955"); force_dense: one-hot mask per row, rows summed into one dense vector */
int cnhe_mat_mul_rowmajor(cnhe_ctx *, const cnhe_vec *const *rows, int n_rows, const cnhe_vec *v, int force_dense, cnhe_vec **out);
/* the same product for a contiguous slice of the rows (rows[i] = global row first_row + i of total_rows): the unit of the
 * intra-inference multi-GPU split of SURVEY.md 8e.  force_dense: one-hot masks at the global columns, so the ranks' partial results add up
 * to the full product ("EncryptedSealBfvMatrix.cs:92-116" sums the masked rows the same way); otherwise this slice's sparse elements. */
int cnhe_mat_mul_rowmajor_shard(cnhe_ctx *, const cnhe_vec *const *rows, int n_rows, const cnhe_vec *v, int force_dense, int first_row,
                                int total_rows, cnhe_vec **out);
/* Whole PoolLayer.Apply with weights ("NeuralNetworks/PoolLayer.cs:149-229"): out[m] = sum_k weights[m][k] * in[gather[m*K+k]]
 * + bias[m].  weights[m] is a plain SPARSE vector of dim K, bias[m] a plain DENSE vector (or NULL); gather < 0 is a
 * padded tap (the reference multiplies a fresh encryption of zero there; we add nothing -- same decryption). */
int cnhe_layer_conv_dense(cnhe_ctx *, const cnhe_vec *const *in, int n_in, const int32_t *gather, const cnhe_vec *const *weights,
                          const cnhe_vec *const *bias, int M, int K, cnhe_vec **out /*M*/);
/* SquareActivation.Apply over a whole matrix ("NeuralNetworks/SquareActivation.cs:10-13"): out[i] = in[i] (.) in[i] */
int cnhe_layer_square(cnhe_ctx *, const cnhe_vec *const *in, int n, cnhe_vec **out /*n*/);

/* ---- micro-benchmark / kernel-level entry points on caller-owned device memory ("raw") --------------------------- */
int cnhe_dev_alloc(cnhe_ctx *, size_t words, uint64_t *dptr);
int cnhe_dev_free(cnhe_ctx *, uint64_t dptr);
int cnhe_dev_upload(cnhe_ctx *, uint64_t dptr, const uint64_t *src, size_t words);
int cnhe_dev_download(cnhe_ctx *, uint64_t *dst, uint64_t dptr, size_t words);
/* n_polys residue polynomials at src; polynomial b uses modulus id mod_base + b % mod_count
 * (ids: see cnhe_context_bsk_moduli) */
int cnhe_raw_ntt(cnhe_ctx *, uint64_t src, uint64_t dst, int n_polys, int mod_base, int mod_count, int inverse);
int cnhe_raw_multiply(cnhe_ctx *, int channel, uint64_t a, uint64_t b, int n, uint64_t out3);     /* Evaluator.Multiply, size 3 out */
int cnhe_raw_relinearize(cnhe_ctx *, int channel, uint64_t in3, int n, uint64_t out2);
int cnhe_raw_multiply_relin(cnhe_ctx *, int channel, uint64_t a, uint64_t b, int n, uint64_t out2);
int cnhe_raw_apply_galois(cnhe_ctx *, int channel, uint64_t in, int n, uint64_t galois_elt, uint64_t out);
int cnhe_raw_rotate_rows(cnhe_ctx *, int channel, uint64_t in, int n, int steps, uint64_t out);
int cnhe_raw_behz_lift(cnhe_ctx *, uint64_t in_cts, int n, uint64_t out_together);
int cnhe_raw_behz_floor(cnhe_ctx *, int channel, uint64_t d_together, int n, uint64_t out3);
int cnhe_dev_copy(cnhe_ctx *, uint64_t dst, uint64_t src, size_t words); /* device to device, on the context stream */
/* per-kernel-family device timing (CUDA events around the launches on the context stream): enable, run, collect.
 * family: 0 ntt forward (incl. digit variant), 1 ntt inverse, 2 behz element-wise, 3 key-switch mac, 4 scalar mac layer, 5 other */
int cnhe_prof_enable(cnhe_ctx *, int on);
int cnhe_prof_collect(cnhe_ctx *, int family, double *total_ms, uint64_t *launches, double *algorithmic_bytes);
int cnhe_raw_event_timing(cnhe_ctx *, int start); /* start=1: record start event; start=0: record stop, return via cnhe_raw_elapsed_ms */
int cnhe_raw_elapsed_ms(cnhe_ctx *, float *ms);
uint64_t cnhe_kernel_launch_count(const cnhe_ctx *); /* kernels launched by this library since context creation */

#ifdef __cplusplus
}
#endif
#endif
