// Host-side number theory for building the device tables of libcnhe (product code; independent of oracle/).
#pragma once
#include <cstdint>
#include <vector>

namespace cnhe {
namespace hm {

typedef unsigned long long u64;
typedef unsigned __int128 u128;

inline u64 mul(u64 a, u64 b, u64 p) { return (u64)((u128)a * b % p); }
inline u64 add(u64 a, u64 b, u64 p) { u64 s = a + b; return (s >= p || s < a) ? s - p : s; }
inline u64 sub(u64 a, u64 b, u64 p) { return a >= b ? a - b : a + p - b; }
inline u64 neg(u64 a, u64 p) { return a ? p - a : 0; }
inline u64 pw(u64 a, u64 e, u64 p) {
    u64 r = 1 % p;
    a %= p;
    for (; e; e >>= 1, a = mul(a, a, p))
        if (e & 1) r = mul(r, a, p);
    return r;
}
inline u64 inv(u64 a, u64 p) { return pw(a, p - 2, p); } // p prime
inline bool is_prime(u64 n) {
    static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (u64 b : bases)
        if (n % b == 0) return n == b;
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) { d >>= 1; ++s; }
    for (u64 b : bases) {
        u64 x = pw(b, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int i = 1; i < s && witness; ++i) {
            x = mul(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}
inline u64 bit_reverse(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < bits; ++i, x >>= 1) r = (r << 1) | (x & 1);
    return r;
}
// smallest primitive `order`-th root of unity modulo prime p (order a power of two dividing p-1)
inline u64 minimal_primitive_root(u64 order, u64 p) {
    u64 cof = (p - 1) / order, g = 0;
    for (u64 x = 2;; ++x) {
        u64 c = pw(x, cof, p);
        if (pw(c, order / 2, p) == p - 1) { g = c; break; }
    }
    u64 step = mul(g, g, p), cur = g, best = g;
    for (u64 i = 1; i < order / 2; ++i) {
        cur = mul(cur, step, p);
        if (cur < best) best = cur;
    }
    return best;
}
inline u64 shoup(u64 w, u64 p) { return (u64)(((u128)w << 64) / p); }
inline void barrett_ratio(u64 p, u64 &r0, u64 &r1) {
    u128 r = (~(u128)0) / p; // p odd, so this equals floor(2^128/p)
    r0 = (u64)r;
    r1 = (u64)(r >> 64);
}
inline int bit_length(u64 v) { return v ? 64 - __builtin_clzll(v) : 0; }
// product of all entries but `skip` (skip < 0: all), modulo p
inline u64 product_mod(const std::vector<u64> &v, int skip, u64 p) {
    u64 r = 1 % p;
    for (int i = 0; i < (int)v.size(); ++i)
        if (i != skip) r = mul(r, v[i] % p, p);
    return r;
}
// floor(prod(v)/d) mod m and prod(v) mod d, by schoolbook long division on 64-bit limbs
inline void div_product(const std::vector<u64> &v, u64 d, std::vector<u64> &quot_limbs, u64 &rem) {
    std::vector<u64> acc{1};
    for (u64 f : v) {
        u64 carry = 0;
        for (auto &limb : acc) {
            u128 x = (u128)limb * f + carry;
            limb = (u64)x;
            carry = (u64)(x >> 64);
        }
        if (carry) acc.push_back(carry);
    }
    quot_limbs.assign(acc.size(), 0);
    u128 r = 0;
    for (size_t i = acc.size(); i-- > 0;) {
        u128 cur = (r << 64) | acc[i];
        quot_limbs[i] = (u64)(cur / d);
        r = cur % d;
    }
    rem = (u64)r;
}
inline u64 limbs_mod(const std::vector<u64> &limbs, u64 m) {
    u128 r = 0;
    for (size_t i = limbs.size(); i-- > 0;) r = ((r << 64) | limbs[i]) % m;
    return (u64)r;
}

} // namespace hm
} // namespace cnhe
