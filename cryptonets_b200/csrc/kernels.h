// Kernel launch interface between the host runtime (context.cu / vec.cu / api.cu) and the CUDA kernels.
// Every launcher is asynchronous on the given stream and returns the cudaError_t of the launch.
#pragma once
#include "modarith.cuh"

namespace cnhe {

// NTT tables of one modulus in HBM (N words each); the kernels index an array of these by modulus id:
//   0..k-1 coefficient primes q_i, k..k+kb-1 the BEHZ base Bsk (aux primes then m_sk), k+kb.. plain moduli.
struct NttTab {
    const u64 *w, *ws;   // psi^bitrev(i) and floor(w 2^64 / p)        (forward, Cooley-Tukey order)
    const u64 *iw, *iws; // psi^-bitrev(i) and its Shoup quotient       (inverse, Gentleman-Sande order)
    u64 inv_n, inv_n_s;  // N^-1 mod p and its Shoup quotient
    DMod mod;
    // FP64 butterfly path (p < 2^50): the same twiddles as exact doubles, centred in (-p/2, p/2]
    const double *wd, *iwd;
    // the 15N/16 twiddles of the four unit-stride stages, transposed for the persistent kernels: wd_hi[m * N/16 + j] is the m-th twiddle
    // (m = 0..14: 1 + 2 + 4 + 8 per stage) of the 16-coefficient group j in the last forward pass; iwd_hi likewise for the first inverse pass
    const double *wd_hi, *iwd_hi;
    double pd, pinv, inv_n_d;
    double inv_n_w_d;                   // iw[1] * N^-1 mod p, centred: the last inverse stage carries the N^-1 scaling
    int fp_ok;                          // 1 when the FP64 path is exact for this modulus and N
    unsigned fwd_recenter, inv_recenter; // forward: bit i = re-centre at the start of pass i; inverse: bit v = re-centre the sums of stage v
    // N = 16384 runs as two half-size transforms on a CTA pair (ntt.cu, "split"): wd_hi / iwd_hi then hold the two halves' twiddle
    // tables ([half][8192], indexed like wd / iwd of an 8192-point transform) and this is the forward schedule of the passes 1+5 | 4 | 4
    unsigned fwd_recenter_split;
    int fwd_out_rc;                     // lazy forward output must be re-centred (its bound squared would overflow the consumer's product)
    double fwd_out_bound;               // |forward lazy output| <= fwd_out_bound * p
};

// Base-2^w digit decomposition used by relinearisation / Galois key switching: digit d comes from residue
// src[d] of the target polynomial, bits [shift[d], shift[d]+w).
struct DigitMap {
    unsigned char src[64];
    unsigned char shift[64];
    int D;
    u64 mask;
};

// `fp` argument of the NTT launchers
enum NttFormat { NTT_FP = 1, NTT_IN_F = 2, NTT_OUT_F = 4 };
enum NttLoad { NTT_LOAD_PLAIN = 0, NTT_LOAD_DIGIT = 1 };
enum NttStore { NTT_STORE_PLAIN = 0, NTT_STORE_ADD = 1 };

// In-place / out-of-place batched negacyclic NTT.  Polynomial b (0 <= b < n_polys) uses modulus
// mod_base + (b % mod_count).  src == dst allowed.
// fp: 1 = every modulus of the range has fp_ok (FP64 butterflies), 0 = integer Harvey butterflies (any p < 2^62)
cudaError_t launch_ntt_forward(const u64 *src, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base, int mod_count, int fp,
                               cudaStream_t s);
cudaError_t launch_ntt_inverse(const u64 *src, u64 *dst, int n_polys, int logn, const NttTab *tabs, int mod_base, int mod_count, int fp,
                               cudaStream_t s);
// dst[((c*k + l)*D + d)] = NTT_l( digit d of target[c] )   target: [n_ct][k][N] coefficient form
// ciphertext c's k-residue target polynomial starts at target + c * ct_stride (words)
cudaError_t launch_ntt_forward_digits(const u64 *target, size_t ct_stride, u64 *dst, int n_ct, int k, const DigitMap &dm, int logn,
                                      const NttTab *tabs, int fp, cudaStream_t s);
// dst[b] = INTT(src[b]) + base[(b / base_group) * base_stride + (b % base_group) * N]  (mod p)
cudaError_t launch_ntt_inverse_add(const u64 *src, const u64 *base, int base_group, size_t base_stride, u64 *dst, int n_polys, int logn,
                                   const NttTab *tabs, int mod_base, int mod_count, int fp, cudaStream_t s);
// pass structure shared by host (re-centring masks) and device: radix (log2) of each pass, forward and inverse
int ntt_pass_radices(int logn, int inverse, int *radices /*4*/);
int ntt_kernel_smem_bytes(int logn);

} // namespace cnhe

// ---------------------------------------------------------------------------------------------------------------
namespace cnhe {

constexpr int KMAX = 9;   // up to 9 coefficient primes (N = 16384 default table)
constexpr int KBMAX = 12; // up to 12 primes in the BEHZ base Bsk (auxiliary primes + m_sk)

// Everything the BEHZ kernels need that depends only on (q, Bsk, m~): SEAL 3.2 util::BaseConverter::generate.
struct BehzConst {
    int k, kb, centered_mtilde, pad_; // k coefficient primes; kb primes in Bsk = (kb-1) auxiliary primes then m_sk
    DMod q[KMAX], bsk[KBMAX];
    u64 inv_qhat_mod_q[KMAX];        // (q/q_i)^-1 mod q_i
    u64 mtilde_inv_qhat_mod_q[KMAX]; // m~ (q/q_i)^-1 mod q_i
    u64 qhat_mod_mtilde[KMAX];       // (q/q_i) mod 2^32
    u64 inv_q_mod_mtilde;            // q^-1 mod 2^32
    u64 qhat_mod_bsk[KBMAX][KMAX];
    u64 q_mod_bsk[KBMAX], inv_q_mod_bsk[KBMAX], inv_mtilde_mod_bsk[KBMAX];
    u64 inv_bhat_mod_b[KBMAX];       // (B/b_j)^-1 mod b_j
    u64 bhat_mod_q[KMAX][KBMAX];     // (B/b_j) mod q_i
    u64 bhat_mod_msk[KBMAX];
    u64 inv_B_mod_msk, B_mod_q[KMAX];
};
// The same constants as exact doubles, centred in (-p/2, p/2], for the FP64 kernels (behz_fp.cu); valid when every
// coefficient and Bsk prime is below 2^50.
struct BehzConstF {
    int k, kb, centered_mtilde, pad_;
    double qd[KMAX], qinv[KMAX], bd[KBMAX], binv[KBMAX];
    double inv_qhat_mod_q[KMAX], mtilde_inv_qhat_mod_q[KMAX];
    u64 qhat_mod_mtilde[KMAX], inv_q_mod_mtilde;
    double qhat_mod_bsk[KBMAX][KMAX];
    double q_mod_bsk[KBMAX], inv_q_mod_bsk[KBMAX], inv_mtilde_mod_bsk[KBMAX];
    double inv_bhat_mod_b[KBMAX], bhat_mod_q[KMAX][KBMAX], bhat_mod_msk[KBMAX];
    double inv_B_mod_msk, msk_half, B_mod_q[KMAX];
    u64 q_u[KMAX], b_u[KBMAX]; // the moduli as integers (sign fix-up of canonical outputs on the integer pipe)
};
// Constants of the folded fast_floor / fastbconv_sk kernel (k_behz_floor_fold_fp), per plaintext modulus: every product by a constant
// that SEAL applies in sequence (x t, x q-hat_i^-1, x q^-1, x B-hat_j^-1) is merged into the next conversion matrix on the host --
// the same residues (modular arithmetic is exact), 19 % fewer FP64 instructions.
struct FloorConstF {
    int k, kb, pad0_, pad1_;
    double qd[KMAX], qinv[KMAX], bd[KBMAX], binv[KBMAX];
    double xq[KMAX];              // t * (q/q_i)^-1 mod q_i
    double xb[KBMAX];             // j < kb-1: t * q^-1 * B-hat_j^-1 mod p_j;  j = kb-1 (m_sk): t * q^-1 mod m_sk
    double conv[KBMAX][KMAX];     // (q/q_i) * q^-1 (* B-hat_j^-1 for j < kb-1) mod p_j
    double bhat_mod_q[KMAX][KBMAX], bhat_mod_msk[KBMAX];
    double inv_B_mod_msk, msk_half, B_mod_q[KMAX];
};
// Per plaintext modulus t.
struct PlainConst {
    u64 t, threshold;                 // (t+1)/2
    u64 delta[KMAX], q_mod_t[KMAX];   // floor(q/t) mod q_i, (q mod t) mod q_i
    // decryption (Decryptor::decrypt, gamma base)
    DMod tmod, gmod;
    u64 tgamma_mod_q[KMAX], qhat_mod_t[KMAX], qhat_mod_gamma[KMAX];
    u64 neg_inv_q_mod_t, neg_inv_q_mod_gamma, inv_gamma_mod_t, gamma;
};

// ---- elementwise over ciphertext words (words = n * size * k * N; residue of word w is (w / N) % k)
cudaError_t launch_ct_add(const u64 *a, const u64 *b, u64 *out, size_t words, int k, int logn, const BehzConst *bc, int sub, cudaStream_t s);
cudaError_t launch_ct_negate(const u64 *a, u64 *out, size_t words, int k, int logn, const BehzConst *bc, cudaStream_t s);
// out = sum_j in_ptrs[j]  (AddMany)
cudaError_t launch_ct_add_many(const u64 *const *in_ptrs, int n_in, u64 *out, size_t words, int k, int logn, const BehzConst *bc, cudaStream_t s);
// ct (size polys) (+/-)= Delta*plain on c0; plain has `coeffs` coefficients mod t.  n cts, plain shared (plain_stride 0) or per ct.
cudaError_t launch_ct_add_plain(const u64 *ct, u64 *out, int n, int size, const u64 *plain, size_t plain_stride, int coeffs, int k,
                                int logn, const BehzConst *bc, PlainConst pc, int sub, cudaStream_t s);
// out[n] = in[n] * scalar (constant plaintext, lifted per residue)      (multiply_plain monomial path, exponent 0)
cudaError_t launch_ct_scale(const u64 *in, u64 *out, int n, int size, const u64 *scalars /*n values mod t, device*/, int k, int logn,
                            const BehzConst *bc, PlainConst pc, cudaStream_t s);
// lifted[l][x] = plain[x] (+ q_l - t if in the upper half)   (multiply_plain generic path)
cudaError_t launch_plain_lift(const u64 *plain, u64 *lifted, int n, int coeffs, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s);
// out[c][part][l][x] = a[c or 0][part][l][x] * b[c or 0][l][x]
cudaError_t launch_dyadic_bcast(const u64 *a, const u64 *b, u64 *out, int n, int size, int a_per_ct, int b_per_ct, int k, int logn,
                                const BehzConst *bc, cudaStream_t s);
// Galois: out[c] = (perm(c0), 0), perm1[c] = perm(c1)   (util::apply_galois)
cudaError_t launch_galois(const u64 *in, u64 *out_base, u64 *perm_c1, int n, u64 elt_inv, int k, int logn, const BehzConst *bc, cudaStream_t s,
                          int add_back = 0); // add_back: base = (c0 + perm(c0), c1) -> key switch + base = x + rotate(x)
// the same with the n input ciphertexts given by a device pointer table
cudaError_t launch_galois_gather(const u64 *const *in_ptrs, u64 *out_base, u64 *perm_c1, int n, u64 elt_inv, int k, int logn, const BehzConst *bc,
                                 cudaStream_t s);

// ---- K4: out[m] = sum_k w[m][k] * in[gather[m][k]] (+ Delta*bias[m] on coefficient 0 of c0)
struct MacTile {
    int n_out;       // outputs in this tile (<= 8) sharing one gather row
    int gather_row;  // row index into gather[] (K entries)
    int out_index[8];
};
// w_ptrs[m]: K weights mod t (device); bias[m] mod t or null
cudaError_t launch_mac_layer(const u64 *const *in_ptrs, const int *gather, const MacTile *tiles, int n_tiles, const u64 *const *w_ptrs,
                             const u64 *bias, int K, u64 *const *out_ptrs, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s);

// small-weight variant: wd[m*K + kk] = signed weight as an exact double (|w| < 2^17 and K*|w|*2^26 < 2^52, host-checked)
cudaError_t launch_mac_layer_fp(const u64 *const *in_ptrs, const int *gather, const MacTile *tiles, int n_tiles, const double *wd, const u64 *bias,
                                int K, u64 *const *out_ptrs, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s);

// ---- K5: BEHZ multiply pieces
// in: ct pointers (each [2][k][N]); out together layout [n][2][k+kb][N] (q residues copied, then Bsk residues)
cudaError_t launch_behz_lift(const u64 *const *ct_ptrs, u64 *out, int n, int logn, const BehzConst *bc, cudaStream_t s);
// d[n][3][2k+1][N] from NTT-form a,b [n][2][2k+1][N]
cudaError_t launch_behz_tensor(const u64 *a, const u64 *b, u64 *d, int n, int kt, int logn, const BehzConst *bc, cudaStream_t s);
// d (coefficient form) -> times t, fast_floor, fastbconv_sk -> out3[n][3][k][N]
cudaError_t launch_behz_floor(const u64 *d, u64 *out3, int n, u64 t, int logn, const BehzConst *bc, cudaStream_t s);
// lazy = 1: the buffers exchanged with the NTT kernels (lift output, tensor input/output, floor input, digit input, accumulator
// output) hold lazy doubles (fparith.cuh) -- pair with NTT_IN_F / NTT_OUT_F on the transforms in between
cudaError_t launch_behz_lift_fp(const u64 *const *ct_ptrs, u64 *out, int n, int logn, const BehzConstF *f, int lazy, cudaStream_t s);
cudaError_t launch_behz_tensor_fp(const u64 *a, const u64 *b, u64 *d, int n, int kt, int logn, const BehzConstF *f, int lazy, cudaStream_t s);
cudaError_t launch_behz_floor_fp(const u64 *d, u64 *out3, int n, u64 t, int logn, const BehzConstF *f, int lazy, cudaStream_t s);
// folded constants + software-pipelined loads (lazy input only)
cudaError_t launch_behz_floor_fold_fp(const u64 *d, u64 *out3, int n, int logn, const FloorConstF *f, cudaStream_t s);
cudaError_t launch_ks_mac_fp(const u64 *digits, const u64 *key, u64 *acc, int n, int D, int k, int logn, const BehzConstF *f, int lazy,
                             cudaStream_t s);
// ---- K6: key-switch inner product. digits [n][D][k][N] (NTT), key [D][2][k][N] (NTT) -> acc [n][2][k][N] (NTT)
cudaError_t launch_ks_mac(const u64 *digits, const u64 *key, u64 *acc, int n, int D, int k, int logn, const BehzConst *bc, cudaStream_t s);
// split a size-3 array [n][3][k][N] view: base[n][2][k][N] = (c0,c1), c2[n][k][N]

// ---- sampling / encode / encrypt / decrypt
enum SampleKind { SAMPLE_TERNARY = 0, SAMPLE_NOISE = 1, SAMPLE_UNIFORM = 2 };
// out[i][l][x] for i<n: stream ids stream0 + i*stream_step (+ l for UNIFORM); lifted into each residue
// Dense layer (every output reads the same K inputs) on the integer tensor cores: wfrag = signed 8-bit weights packed as m16n8k32
// A fragments [ceil(M/16)][ceil(K/32)][32 lanes][16 bytes]; weights in [-254, 254] are split W = W1 + W2 (wfrag2, may be null);
// limbs = ceil(bits(q)/8); needs K*254*255 < 2^31
cudaError_t launch_mac_dense_imma(const u64 *const *in_ptrs, const void *wfrag, const void *wfrag2, const u64 *bias, int K, int M, int limbs,
                                  u64 *const *out_ptrs, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s);
// Scalar-MAC layers on tcgen05 (mac_umma.cu).  The layer's inputs are rows of one slab (input i at slab + i * slab_stride_words); a
// BUNDLE is up to 128 outputs whose taps lie in a window of consecutive inputs: chunk j of the bundle multiplies the 32 inputs that start
// at row chunk_rows[chunk0 + j] (bit 30 set: a row of the scratch slab that holds the W2 taps) with the 128 x 32 weight block at
// wpack + a_off + j * 4096.  Outputs and constant biases are listed in bundle order (out0 = first entry of the bundle).
struct UmBundle {
    int chunk0, n_chunks, a_off, n_out, out0;
};
struct UmmaLaunch {
    const u64 *slab;            // input 0
    size_t slab_stride_words;   // distance between consecutive inputs
    size_t slab_rows;           // number of inputs
    const u64 *scratch;         // W2 taps gathered side by side (ct_words apart), or null
    size_t scratch_rows;
    const UmBundle *bundles;    // device
    int n_bundles;
    const int *chunk_rows;      // device
    int total_chunks;           // chunks per tile (sum over the bundles)
    const unsigned char *wpack; // device: a_bytes of weight blocks (identical matrices stored once)
    int a_bytes;
    u64 *const *out_ptrs;       // device, bundle order
    const u64 *bias;            // device, bundle order; null = no constant bias
    int n_out_total;
    int limbs, k, logn;
    const BehzConst *bc;
    PlainConst pc;
};
bool mac_umma_fits(int a_bytes, int total_chunks, int n_out_total, int n_bundles, int limbs);
void mac_umma_pack(const signed char *w, int rows, int cols, unsigned char *out); // cols a multiple of 32; out: cols / 32 * 4096 bytes
cudaError_t launch_mac_umma(const UmmaLaunch &a, cudaStream_t s);
// 2-D tensor map over 64-bit words (ntt.cu): rows of inner_words words, row_stride bytes apart; swizzle128: rows of the box are 128 bytes
cudaError_t make_word_map_2d(void *map, const u64 *base, size_t inner_words, size_t rows, size_t row_stride, unsigned box_words, unsigned box_rows,
                             int swizzle128);
cudaError_t launch_sample(u64 *out, int n, int kind, const RngKey &seed, u64 stream0, u64 stream_step, int k, int logn, const BehzConst *bc, cudaStream_t s);
// plain[i][index_map[j]] = values[i][j]  (j < count), zero elsewhere
cudaError_t launch_encode_scatter(const u64 *values, u64 *plain, int n, int count, const u32 *index_map, int logn, cudaStream_t s);
// plain[i][index_map[first_col + i]] = 1, zero elsewhere (one-hot slot masks, encoded form before the inverse transform)
cudaError_t launch_onehot_scatter(u64 *plain, int n, int first_col, const u32 *index_map, int logn, cudaStream_t s);
cudaError_t launch_decode_gather(const u64 *plain_ntt, u64 *values, int n, const u32 *index_map, int logn, cudaStream_t s);
// ct[i] (already u*pk, coefficient form) += (e0 + Delta*m_i, e1)
cudaError_t launch_encrypt_finish(u64 *ct, const u64 *plain, size_t plain_stride, int n, int coeffs, const RngKey &seed, u64 nonce0, int k, int logn,
                                  const BehzConst *bc, PlainConst pc, cudaStream_t s);
// x[n][k][N] = c0 + c1*s (coefficient form) -> plain[n][N]
cudaError_t launch_decrypt_round(const u64 *x, u64 *plain, int n, int k, int logn, const BehzConst *bc, PlainConst pc, cudaStream_t s);
cudaError_t launch_fill_zero(u64 *p, size_t words, cudaStream_t s);
// keys[d][0][src[d]][x] += factors[d] * target[src[d]][x]   (KeyGenerator: the 2^{jw} s' term of key-switching key d)
cudaError_t launch_key_add_scaled(u64 *keys, const u64 *target, const u64 *factors, const DigitMap &dm, int k, int logn, const BehzConst *bc,
                                  cudaStream_t s);

} // namespace cnhe
