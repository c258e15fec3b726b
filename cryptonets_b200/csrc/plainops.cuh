// Plaintext scaling shared by the element-wise kernels (poly_ops.cu) and the tensor-core dense layer (mac_imma.cu).
#pragma once
#include "kernels.h"
#include "modarith.cuh"

namespace cnhe {

// Delta*m (+ the upper-half increment q mod t for "negative" m): SEAL 3.2 Encryptor::preencrypt / Evaluator::add_plain
__device__ __forceinline__ u64 scale_plain(u64 m, int l, const DMod &q, const PlainConst &pc) {
    U128 v = mul64wide(pc.delta[l], m);
    if (m >= pc.threshold) add128(v, pc.q_mod_t[l]);
    return barrett128(v, q);
}

} // namespace cnhe
