// Internal header shared by the translation units that implement the C ABI (vec.cu: vectors, layers; wire.cu: wire formats).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "../../include/cnhe.h"
#include "runtime.h"

using namespace cnhe;

struct cnhe_ctx {
    Context *c;
};

struct cnhe_vec {
    Context *ctx = nullptr;
    uint64_t dim = 0;
    double scale = 1.0;
    int format = CNHE_DENSE;
    bool enc = false;
    int blocks = 0; // ciphertexts / plaintexts per channel
    std::vector<BufRef> buf; // per channel: enc -> blocks*2kN words, plain dense -> blocks*N words, plain sparse -> `blocks` scalars
    std::vector<size_t> off;
    std::vector<std::vector<u64>> scalars; // plain sparse: host copy of the constants (mod t)
    bool is_const = false;                 // plain dense whose every plaintext is a constant polynomial
    std::vector<u64> const_val;            // per channel constant (mod t) when is_const

    u64 *ptr(int ch) const { return buf[ch]->p + off[ch]; }
    size_t unit() const { return enc ? ctx->ct_words() : (format == CNHE_DENSE ? (size_t)ctx->N : 1); }
    u64 *block(int ch, int b) const { return ptr(ch) + (size_t)b * unit(); }
};

int set_err(int code, const std::string &m); // thread-local last error (vec.cu)

#define API_BEGIN(CTX)                                                                                                 \
    if (!(CTX)) return set_err(CNHE_ERR_INVALID, "null context");                                                      \
    Context &c = *(CTX)->c;                                                                                            \
    try {                                                                                                              \
        std::lock_guard<std::recursive_mutex> lock(c.mu);                                                              \
        CNHE_CUDA(cudaSetDevice(c.device));                                                                            \
        c.set_channel(0);                                                                                              \
        ws_release_all(c);
#define API_END                                                                                                        \
    }                                                                                                                  \
    catch (const Error &e) { return set_err(e.code, e.what()); }                                                      \
    catch (const std::exception &e) { return set_err(CNHE_ERR_INVALID, e.what()); }                                   \
    return CNHE_OK;
static inline void fail(const char *m) { throw Error(CNHE_ERR_INVALID, m); }
cnhe_vec *new_vec(Context &c, uint64_t dim, double scale, int format, bool enc, int blocks);
void alloc_channels(cnhe_vec *v);
void same_ctx(Context &c, const cnhe_vec *v);

