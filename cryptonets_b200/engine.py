"""Thin object layer over the C ABI: Engine (one cnhe_ctx) and Vec (one cnhe_vec handle).

This is binding code only -- every method is one call into libcnhe.so.  The reference-shaped API (IFactory, IVector,
IMatrix, layers) lives in he.py / layers.py on top of this."""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import U64P, DBLP, VECP, check

DENSE, SPARSE = 0, 1
ALL_SLOTS = 0x7FFFFFFF


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a):
    return a.ctypes.data_as(U64P)


def _vec_array(vecs):
    arr = (VECP * len(vecs))()
    for i, v in enumerate(vecs):
        arr[i] = None if v is None else v.h
    return arr


class Vec:
    """Owning handle of a cnhe_vec (== EncryptedSealBfvVector).  Dispose() mirrors IDisposable."""

    __slots__ = ("eng", "h", "_m", "__weakref__")

    def __init__(self, eng, handle):
        self.eng = eng
        self.h = VECP(handle) if not isinstance(handle, VECP) else handle
        self._m = None  # metadata cache: a cnhe_vec is immutable except for RegisterScale / RegisterDim
        eng._live.add(self)

    def dispose(self):
        if self.h:
            if self.eng.h:  # a closed engine has already released every vector
                self.eng.L.cnhe_vec_destroy(self.h)
            self.h = VECP(None)

    def __del__(self):
        try:
            self.dispose()
        except Exception:
            pass

    def meta(self):
        if self._m is not None:
            return self._m
        dim, bs = C.c_uint64(), C.c_uint64()
        scale = C.c_double()
        fmt, enc, blocks = C.c_int(), C.c_int(), C.c_int()
        check(self.eng.L.cnhe_vec_meta(self.h, C.byref(dim), C.byref(scale), C.byref(fmt), C.byref(enc), C.byref(blocks), C.byref(bs)))
        self._m = dict(dim=dim.value, scale=scale.value, format=fmt.value, encrypted=bool(enc.value), blocks=blocks.value, block_size=bs.value)
        return self._m

    dim = property(lambda s: s.meta()["dim"])
    scale = property(lambda s: s.meta()["scale"])
    format = property(lambda s: s.meta()["format"])
    is_encrypted = property(lambda s: s.meta()["encrypted"])
    blocks = property(lambda s: s.meta()["blocks"])

    def register_scale(self, scale):
        check(self.eng.L.cnhe_vec_register_scale(self.h, float(scale)))
        self._m = None

    def register_dim(self, dim):
        check(self.eng.L.cnhe_vec_register_dim(self.h, int(dim)))
        self._m = None

    def export_raw(self, channel=0, block=0):
        out = np.zeros(self.eng.ct_words, np.uint64)
        check(self.eng.L.cnhe_vec_export_raw(self.eng.h, self.h, channel, block, _p(out), out.size))
        return out

    def device_ptr(self, channel=0):
        p, w = C.c_uint64(), C.c_size_t()
        check(self.eng.L.cnhe_vec_device_ptr(self.h, channel, C.byref(p), C.byref(w)))
        return p.value, w.value


class Engine:
    """One cnhe_ctx: parameters, device tables and keys for P plaintext moduli (== EncryptedSealBfvFactory)."""

    def __init__(self, plain_primes, N=0, dbc_relin=10, dbc_galois=20, small_modulus_count=-1, device=0, coeff_moduli=None, archive=None):
        """archive: bytes of a key archive (cnhe_keys_save / EncryptedSealBfvEnvironment.Save): parameters and keys come from it."""
        self.L = _lib.lib()
        self._live = weakref.WeakSet()
        h = C.c_void_p()
        if archive is not None:
            buf = (C.c_ubyte * len(archive)).from_buffer_copy(archive)
            check(self.L.cnhe_context_load(buf, len(archive), device, C.byref(h)))
            self.h = h
            P = C.c_int()
            check(self.L.cnhe_context_info(self.h, None, None, C.byref(P), None, None, None))
            pp = np.zeros(P.value, np.uint64)
            check(self.L.cnhe_context_plain_moduli(self.h, _p(pp)))
        else:
            pp = _u64(plain_primes)
        if archive is not None:
            pass
        elif coeff_moduli is None:
            check(self.L.cnhe_context_create(_p(pp), len(pp), N, dbc_relin, dbc_galois, small_modulus_count, device, C.byref(h)))
        else:
            cm = _u64(coeff_moduli)
            check(self.L.cnhe_context_create_custom(_p(pp), len(pp), N, _p(cm), len(cm), dbc_relin, dbc_galois, device, C.byref(h)))
        self.h = h
        n, k, P, rd, gd, ge = C.c_uint32(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        check(self.L.cnhe_context_info(self.h, C.byref(n), C.byref(k), C.byref(P), C.byref(rd), C.byref(gd), C.byref(ge)))
        self.N, self.k, self.P = n.value, k.value, P.value
        self.relin_digits, self.galois_digits, self.n_galois = rd.value, gd.value, ge.value
        self.ct_words = 2 * self.k * self.N
        q = np.zeros(self.k, np.uint64)
        check(self.L.cnhe_context_coeff_moduli(self.h, _p(q)))
        self.q = [int(x) for x in q]
        self.primes = [int(x) for x in pp]
        cnt = C.c_int()
        check(self.L.cnhe_context_bsk_moduli(self.h, None, C.byref(cnt)))
        b = np.zeros(cnt.value, np.uint64)
        check(self.L.cnhe_context_bsk_moduli(self.h, _p(b), C.byref(cnt)))
        self.bsk = [int(x) for x in b]
        self.kb = cnt.value
        self.plain_mod_id = self.k + self.kb  # NTT table id of plaintext modulus 0

    def close(self):
        if self.h:
            for v in list(self._live):  # vectors hold device buffers of this context: release them first
                v.dispose()
            self.L.cnhe_context_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- context / keys
    def set_option(self, name, value):
        check(self.L.cnhe_context_set_option(self.h, name.encode(), int(value)))

    def sync(self):
        check(self.L.cnhe_context_sync(self.h))

    def stream(self, channel=0):
        """cudaStream_t (as an integer) of a channel: wrap it with torch.cuda.ExternalStream to order torch / NCCL work with the library's."""
        s = C.c_uint64()
        check(self.L.cnhe_context_stream(self.h, int(channel), C.byref(s)))
        return s.value

    def join_streams(self):
        check(self.L.cnhe_context_join_streams(self.h))

    def fork_streams(self):
        check(self.L.cnhe_context_fork_streams(self.h))

    def keygen(self, seed=None):
        """seed=None: keys and all later encryption randomness from the OS CSPRNG (production).  An integer seed selects the deterministic
        sampler shared with the CPU oracle -- tests only, the keys are predictable."""
        if seed is None:
            check(self.L.cnhe_keys_generate_secure(self.h))
        else:
            check(self.L.cnhe_keys_generate(self.h, int(seed)))

    OP_NAMES = None

    def op_counts(self, reset=False):
        """Evaluator-level operation counters (the reference's OperationsCount)."""
        n = 14
        a = np.zeros(n, np.uint64)
        check(self.L.cnhe_op_counts(self.h, _p(a), n, int(reset)))
        if Engine.OP_NAMES is None:
            Engine.OP_NAMES = [self.L.cnhe_op_name(i).decode() for i in range(n)]
        return {nm: int(v) for nm, v in zip(Engine.OP_NAMES, a)}

    def trace_noise(self, on=True):
        self.set_option("trace_noise", 1 if on else 0)

    def trace_read(self, clear=True):
        """[(operation name, channel, count, budget of the first output, budgets of its first two inputs (-1 unknown), aux)] since the trace
        was last cleared (see cnhe_trace_read in include/cnhe.h)."""
        n = C.c_size_t()
        check(self.L.cnhe_trace_read(self.h, None, 0, C.byref(n), 0))
        a = np.zeros((max(n.value, 1), 8), np.int32)
        check(self.L.cnhe_trace_read(self.h, a.ctypes.data_as(C.POINTER(C.c_int32)), n.value, C.byref(n), int(clear)))
        self.op_counts()
        return [(Engine.OP_NAMES[r[0]], int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]), r[6] / 1000.0) for r in a[:n.value]]

    def save_keys(self, with_private_keys=False):
        """The key archive of EncryptedSealBfvEnvironment.Save as bytes."""
        n = C.c_size_t()
        check(self.L.cnhe_keys_save(self.h, int(with_private_keys), None, 0, C.byref(n)))
        buf = (C.c_ubyte * n.value)()
        check(self.L.cnhe_keys_save(self.h, int(with_private_keys), buf, n.value, C.byref(n)))
        return bytes(buf)

    def write_vector(self, vec):
        """EncryptedSealBfvVector.Write: the text form of one vector."""
        n = C.c_size_t()
        check(self.L.cnhe_vec_write(self.h, vec.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        check(self.L.cnhe_vec_write(self.h, vec.h, buf, n.value, C.byref(n)))
        return buf.raw[: n.value].decode("ascii")

    def read_vector(self, text):
        """EncryptedSealBfvVector.Read; returns (Vec, characters consumed)."""
        raw = text.encode("ascii") if isinstance(text, str) else text
        out, used = VECP(), C.c_size_t()
        check(self.L.cnhe_vec_read(self.h, raw, len(raw), C.byref(out), C.byref(used)))
        return Vec(self, out), used.value

    def galois_elts(self):
        a = np.zeros(self.n_galois, np.uint64)
        check(self.L.cnhe_context_galois_elts(self.h, _p(a)))
        return [int(x) for x in a]

    def _key_words(self, what):
        kN = self.k * self.N
        return {0: kN, 1: 2 * kN, 2: self.relin_digits * 2 * kN, 3: self.galois_digits * 2 * kN}[what]

    def export_key(self, channel, what, arg=0):
        a = np.zeros(self._key_words(what), np.uint64)
        check(self.L.cnhe_keys_export(self.h, channel, what, int(arg), _p(a), a.size))
        return a

    def import_key(self, channel, what, data, arg=0):
        a = _u64(data).ravel()
        check(self.L.cnhe_keys_import(self.h, channel, what, int(arg), _p(a), a.size))

    def set_seed(self, channel, seed):
        check(self.L.cnhe_keys_set_seed(self.h, channel, int(seed)))

    def launch_count(self):
        return int(self.L.cnhe_kernel_launch_count(self.h))

    # ---- vectors
    def _new(self, fn, v, scale, fmt):
        a = np.ascontiguousarray(v, dtype=np.float64).ravel()
        out = VECP()
        check(fn(self.h, a.ctypes.data_as(DBLP), a.size, float(scale), fmt, C.byref(out)))
        return Vec(self, out)

    def encrypt(self, v, scale=1.0, fmt=DENSE):
        return self._new(self.L.cnhe_vec_encrypt, v, scale, fmt)

    def plain(self, v, scale=1.0, fmt=DENSE):
        return self._new(self.L.cnhe_vec_plain, v, scale, fmt)

    def encrypt_many(self, rows, scale=1.0):
        a = np.ascontiguousarray(rows, dtype=np.float64)
        n, dim = a.shape
        out = (VECP * n)()
        check(self.L.cnhe_vecs_encrypt(self.h, a.ctypes.data_as(DBLP), n, dim, float(scale), out))
        return [Vec(self, out[i]) for i in range(n)]

    def decrypt(self, vec):
        out = np.zeros(vec.dim, np.float64)
        check(self.L.cnhe_vec_decrypt(self.h, vec.h, out.ctypes.data_as(DBLP), out.size))
        return out

    def decrypt_residues(self, vec):
        """Per-channel residues [P][dim] of the decryption (DecryptFullPrecision joins them with big integers)."""
        out = np.zeros((self.P, vec.dim), np.uint64)
        check(self.L.cnhe_vec_decrypt_residues(self.h, vec.h, _p(out), out.size))
        return out

    def from_residues(self, residues, scale=1.0, fmt=DENSE, encrypt=True):
        r = _u64(residues).reshape(self.P, -1)
        out = VECP()
        check(self.L.cnhe_vec_from_residues(self.h, _p(r), r.shape[1], float(scale), fmt, int(encrypt), C.byref(out)))
        return Vec(self, out)

    def dispose_many(self, vecs):
        """Release a list of Vec handles with one ABI call."""
        live = [v for v in vecs if v is not None and v.h]
        if live and self.h:
            check(self.L.cnhe_vecs_destroy(_vec_array(live), len(live)))
        for v in live:
            v.h = VECP(None)

    def decrypt_many(self, vecs):
        dim = vecs[0].dim
        out = np.zeros((len(vecs), dim), np.float64)
        check(self.L.cnhe_vecs_decrypt(self.h, _vec_array(vecs), len(vecs), out.ctypes.data_as(DBLP), dim))
        return out

    def import_raw(self, data, blocks, dim, scale=1.0, fmt=DENSE):
        a = _u64(data).ravel()
        assert a.size == self.P * blocks * self.ct_words
        out = VECP()
        check(self.L.cnhe_vec_import_raw(self.h, _p(a), blocks, int(dim), float(scale), fmt, C.byref(out)))
        return Vec(self, out)

    def import_raw_ptr(self, ptr, blocks, dim, scale=1.0, fmt=DENSE):
        """One vector from raw words at a host OR device address, layout [P][blocks][2kN] (multi-GPU exchanges hand over device buffers)."""
        out = VECP()
        check(self.L.cnhe_vec_import_raw(self.h, C.cast(int(ptr), U64P), int(blocks), int(dim), float(scale), fmt, C.byref(out)))
        return Vec(self, out)

    def import_raw_many(self, host_ptr_or_array, n, blocks, dim, scale=1.0, fmt=DENSE):
        """host layout [P][n][blocks][2kN]; accepts a numpy array or a raw host address (e.g. a pinned torch tensor's data_ptr())."""
        if isinstance(host_ptr_or_array, int):
            src = C.cast(host_ptr_or_array, U64P)
        else:
            a = _u64(host_ptr_or_array).ravel()
            assert a.size == self.P * n * blocks * self.ct_words
            src = _p(a)
        out = (VECP * n)()
        check(self.L.cnhe_vecs_import_raw(self.h, src, n, blocks, int(dim), float(scale), fmt, out))
        return self._wrap_many(out, n)

    def export_raw_many(self, vecs, host_ptr=None):
        n, blocks = len(vecs), vecs[0].blocks
        words = self.P * n * blocks * self.ct_words
        if host_ptr is None:
            a = np.zeros(words, np.uint64)
            check(self.L.cnhe_vecs_export_raw(self.h, _vec_array(vecs), n, _p(a), words))
            return a.reshape(self.P, n, blocks, self.ct_words)
        check(self.L.cnhe_vecs_export_raw(self.h, _vec_array(vecs), n, C.cast(host_ptr, U64P), words))
        return None

    def export_raw_many_async(self, vecs, host_ptr):
        """Queue the device-to-host copies behind the producing kernels; returns a ticket for export_wait().  `host_ptr`: pinned memory."""
        n, blocks = len(vecs), vecs[0].blocks
        t = C.c_int()
        check(self.L.cnhe_vecs_export_raw_async(self.h, _vec_array(vecs), n, C.cast(host_ptr, U64P), self.P * n * blocks * self.ct_words, C.byref(t)))
        return t.value

    def export_wait(self, ticket):
        check(self.L.cnhe_export_wait(self.h, int(ticket)))

    def dev_copy(self, dst, src, words):
        check(self.L.cnhe_dev_copy(self.h, int(dst), int(src), int(words)))

    def prof_enable(self, on=True):
        check(self.L.cnhe_prof_enable(self.h, int(on)))

    def prof_collect(self):
        names = ["ntt_forward", "ntt_inverse", "behz_elementwise", "keyswitch_mac", "scalar_mac_layer", "other"]
        out = {}
        for i, nm in enumerate(names):
            ms, n, b = C.c_double(), C.c_uint64(), C.c_double()
            check(self.L.cnhe_prof_collect(self.h, i, C.byref(ms), C.byref(n), C.byref(b)))
            out[nm] = dict(ms=ms.value, launches=n.value, bytes=b.value)
        return out

    def copy(self, vec):
        out = VECP()
        check(self.L.cnhe_vec_copy(self.h, vec.h, C.byref(out)))
        return Vec(self, out)

    def noise_budget(self, vec, channel=0, block=0):
        b = C.c_int()
        check(self.L.cnhe_noise_budget(self.h, vec.h, channel, block, C.byref(b)))
        return b.value

    def _bin(self, fn, a, b):
        out = VECP()
        check(fn(self.h, a.h, b.h, C.byref(out)))
        return Vec(self, out)

    def add(self, a, b):
        return self._bin(self.L.cnhe_vec_add, a, b)

    def sub(self, a, b):
        return self._bin(self.L.cnhe_vec_sub, a, b)

    def pointwise_multiply(self, a, b):
        return self._bin(self.L.cnhe_vec_pointwise_multiply, a, b)

    def sum_all_slots(self, a, length=ALL_SLOTS, force_column=-1):
        out = VECP()
        check(self.L.cnhe_vec_sum_all_slots(self.h, a.h, int(length), int(force_column), C.byref(out)))
        return Vec(self, out)

    def dot_product(self, a, b, length=ALL_SLOTS, force_column=-1):
        out = VECP()
        check(self.L.cnhe_vec_dot_product(self.h, a.h, b.h, int(length), int(force_column), C.byref(out)))
        return Vec(self, out)

    def rotate(self, a, amount):
        out = VECP()
        check(self.L.cnhe_vec_rotate(self.h, a.h, int(amount), C.byref(out)))
        return Vec(self, out)

    def duplicate(self, a, count):
        out = VECP()
        check(self.L.cnhe_vec_duplicate(self.h, a.h, int(count), C.byref(out)))
        return Vec(self, out)

    def permute(self, a, selections, shifts, output_dim):
        sh = (C.c_int * len(shifts))(*[int(s) for s in shifts])
        out = VECP()
        check(self.L.cnhe_vec_permute(self.h, a.h, _vec_array(selections), sh, len(shifts), int(output_dim), C.byref(out)))
        return Vec(self, out)

    def interleave(self, vecs, shift):
        out = VECP()
        check(self.L.cnhe_vecs_interleave(self.h, _vec_array(vecs), len(vecs), int(shift), C.byref(out)))
        return Vec(self, out)

    def stack(self, vecs):
        out = VECP()
        check(self.L.cnhe_vecs_stack(self.h, _vec_array(vecs), len(vecs), C.byref(out)))
        return Vec(self, out)

    def generate_sparse_of_array(self, vecs):
        out = VECP()
        check(self.L.cnhe_vecs_generate_sparse_of_array(self.h, _vec_array(vecs), len(vecs), C.byref(out)))
        return Vec(self, out)

    def mat_mul_colmajor_sparse(self, cols, sparse):
        out = VECP()
        check(self.L.cnhe_mat_mul_colmajor_sparse(self.h, _vec_array(cols), len(cols), sparse.h, C.byref(out)))
        return Vec(self, out)

    def mat_mul_rowmajor(self, rows, v, force_dense=False):
        out = VECP()
        check(self.L.cnhe_mat_mul_rowmajor(self.h, _vec_array(rows), len(rows), v.h, int(force_dense), C.byref(out)))
        return Vec(self, out)

    def mat_mul_rowmajor_shard(self, rows, v, force_dense, first_row, total_rows):
        out = VECP()
        check(self.L.cnhe_mat_mul_rowmajor_shard(self.h, _vec_array(rows), len(rows), v.h, int(force_dense), int(first_row), int(total_rows), C.byref(out)))
        return Vec(self, out)

    def layer_conv_dense(self, inputs, gather, weights, bias, M, K):
        g = None
        if gather is not None:
            g = np.ascontiguousarray(gather, dtype=np.int32)
            assert g.size == M * K
        out = (VECP * M)()
        check(self.L.cnhe_layer_conv_dense(
            self.h, _vec_array(inputs), len(inputs), None if g is None else g.ctypes.data_as(C.POINTER(C.c_int32)), _vec_array(weights),
            None if bias is None else _vec_array(bias), M, K, out))
        return self._wrap_many(out, M)

    def _wrap_many(self, out, n):
        """Vec handles for the n outputs of one batched call: they share dim/scale/format/blocks, so the metadata is queried once."""
        vecs = [Vec(self, out[i]) for i in range(n)]
        if n > 1:
            m = vecs[0].meta()
            for v in vecs[1:]:
                v._m = m
        return vecs

    def layer_square(self, inputs):
        n = len(inputs)
        out = (VECP * n)()
        check(self.L.cnhe_layer_square(self.h, _vec_array(inputs), n, out))
        return self._wrap_many(out, n)

    # ---- raw device arrays (micro-benchmarks, kernel parity tests)
    def dev_alloc(self, words):
        p = C.c_uint64()
        check(self.L.cnhe_dev_alloc(self.h, int(words), C.byref(p)))
        return p.value

    def dev_free(self, ptr):
        check(self.L.cnhe_dev_free(self.h, int(ptr)))

    def dev_upload(self, ptr, data):
        a = _u64(data).ravel()
        check(self.L.cnhe_dev_upload(self.h, int(ptr), _p(a), a.size))

    def dev_download(self, ptr, words):
        a = np.zeros(int(words), np.uint64)
        check(self.L.cnhe_dev_download(self.h, _p(a), int(ptr), a.size))
        return a

    def dev_from(self, data):
        a = _u64(data).ravel()
        p = self.dev_alloc(a.size)
        self.dev_upload(p, a)
        return p

    def raw_ntt(self, src, dst, n_polys, mod_base, mod_count, inverse=False):
        check(self.L.cnhe_raw_ntt(self.h, int(src), int(dst), int(n_polys), int(mod_base), int(mod_count), int(inverse)))

    def raw_multiply(self, ch, a, b, n, out3):
        check(self.L.cnhe_raw_multiply(self.h, ch, int(a), int(b), int(n), int(out3)))

    def raw_relinearize(self, ch, in3, n, out2):
        check(self.L.cnhe_raw_relinearize(self.h, ch, int(in3), int(n), int(out2)))

    def raw_multiply_relin(self, ch, a, b, n, out2):
        check(self.L.cnhe_raw_multiply_relin(self.h, ch, int(a), int(b), int(n), int(out2)))

    def raw_apply_galois(self, ch, src, n, elt, out):
        check(self.L.cnhe_raw_apply_galois(self.h, ch, int(src), int(n), int(elt), int(out)))

    def raw_rotate_rows(self, ch, src, n, steps, out):
        check(self.L.cnhe_raw_rotate_rows(self.h, ch, int(src), int(n), int(steps), int(out)))

    def raw_behz_lift(self, src, n, out):
        check(self.L.cnhe_raw_behz_lift(self.h, int(src), int(n), int(out)))

    def raw_behz_floor(self, ch, d, n, out3):
        check(self.L.cnhe_raw_behz_floor(self.h, ch, int(d), int(n), int(out3)))

    def timer_start(self):
        check(self.L.cnhe_raw_event_timing(self.h, 1))

    def timer_stop_ms(self):
        check(self.L.cnhe_raw_event_timing(self.h, 0))
        ms = C.c_float()
        check(self.L.cnhe_raw_elapsed_ms(self.h, C.byref(ms)))
        return ms.value
