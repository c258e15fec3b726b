"""Wire / on-disk formats of the reference, restated independently in Python (TEST INFRASTRUCTURE: the checker of the product's
C++ implementation in cryptonets_b200/csrc/wire.cu; the product never imports this file).

PARITY WITH THE REAL SEAL 3.2 BINARY STREAMS IS UNPINNED: the container formats below are fully specified by the reference's own C#
(cited), but the SEAL 3.2 `Save` layouts inside them are restated from knowledge of SEAL 3.2.x -- SEAL is not in the reference tree and
cannot be built here.  Each recalled detail is marked [EXT].

Containers (reference C#):
  key archive  `EncryptedSealBfvEnvironment.Save`  ("HE Wrapper/EncryptedSealBfvVector.cs:104-134"): a ZIP archive, one STORED entry
               `environmentNNN` per plaintext modulus = `AtomicSealBfvEncryptedEnvironment.SaveToStream`
               ("HE Wrapper/AtomicSealBfvVector.cs:93-104"): EncryptionParameters, PublicKey, RelinKeys, GaloisKeys, SecretKey
               (an empty SecretKey when saved without private keys).
  vector text  `EncryptedSealBfvVector.Write` (":414-439") wrapping `AtomicSealBfvEncryptedVector.Write`
               ("AtomicSealBfvVector.cs:1273-1303"): text lines, the SEAL objects of one channel concatenated and base64-encoded on one line.
  matrix text  `EncryptedSealBfvMatrix.Write` ("HE Wrapper/EncryptedSealBfvMatrix.cs:182-208").

SEAL 3.2 streams [EXT] (little endian):
  EncryptionParameters::Save   u8 scheme (1 = BFV) | u64 poly_modulus_degree | u64 coeff_mod_count | u64 value per coefficient modulus |
                               u64 plain modulus | f64 noise_standard_deviation
  parms_id                     SHA3-256 over the u64 array [scheme, N, q_0.., t, bits(noise_standard_deviation)], stored as 4 x u64
  Ciphertext::save             parms_id (32 B) | u8 is_ntt_form | u64 size | u64 poly_modulus_degree | u64 coeff_mod_count | f64 scale |
                               IntArray: u64 word count, words
  Plaintext::save              parms_id (32 B; zero for a BFV message plaintext) | f64 scale | IntArray: u64 coefficient count, words
  PublicKey::save              its ciphertext (NTT form)          SecretKey::save   its plaintext (NTT form, k*N words, key parms_id)
  RelinKeys::save              parms_id | i32 decomposition_bit_count | u64 dim1 | per entry: u64 dim2, ciphertexts (NTT form, size 2)
  GaloisKeys::save             same; dim1 = N, entry (galois_elt - 1) / 2, empty entries have dim2 = 0"""
import base64
import hashlib
import io
import struct
import zipfile

import numpy as np

NOISE_STANDARD_DEVIATION = 3.20  # [EXT] util::global_variables::default_noise_standard_deviation of SEAL 3.2


def parms_id(N, q, t):
    data = [1, N] + [int(x) for x in q] + [int(t)]
    raw = struct.pack("<%dQ" % len(data), *data) + struct.pack("<d", NOISE_STANDARD_DEVIATION)
    return hashlib.sha3_256(raw).digest()


def save_parms(N, q, t):
    return struct.pack("<BQQ", 1, N, len(q)) + struct.pack("<%dQ" % len(q), *[int(x) for x in q]) + struct.pack("<Qd", int(t), NOISE_STANDARD_DEVIATION)


def load_parms(f):
    scheme, N, k = struct.unpack("<BQQ", f.read(17))
    assert scheme == 1
    q = list(struct.unpack("<%dQ" % k, f.read(8 * k)))
    t, sd = struct.unpack("<Qd", f.read(16))
    return N, q, t, sd


def _int_array(words):
    w = np.ascontiguousarray(words, dtype="<u8").ravel()
    return struct.pack("<Q", w.size) + w.tobytes()


def _read_int_array(f):
    (n,) = struct.unpack("<Q", f.read(8))
    return np.frombuffer(f.read(8 * n), dtype="<u8").astype(np.uint64)


def save_ciphertext(pid, words, N, k, size=2, ntt=False, scale=1.0):
    return pid + struct.pack("<BQQQd", 1 if ntt else 0, size, N, k, scale) + _int_array(words)


def load_ciphertext(f):
    pid = f.read(32)
    ntt, size, N, k, scale = struct.unpack("<BQQQd", f.read(33))
    words = _read_int_array(f)
    assert words.size == size * N * k
    return dict(parms_id=pid, ntt=bool(ntt), size=size, N=N, k=k, scale=scale, words=words)


def save_plaintext(words, pid=b"\0" * 32, scale=1.0):
    return pid + struct.pack("<d", scale) + _int_array(words)


def load_plaintext(f):
    pid = f.read(32)
    (scale,) = struct.unpack("<d", f.read(8))
    return dict(parms_id=pid, scale=scale, words=_read_int_array(f))


def save_kswitch(pid, dbc, entries, N, k):
    """entries: list (dim1) of None or arrays [D][2kN]"""
    out = [pid, struct.pack("<iQ", dbc, len(entries))]
    for e in entries:
        if e is None:
            out.append(struct.pack("<Q", 0))
            continue
        e = np.asarray(e, dtype=np.uint64).reshape(-1, 2 * k * N)
        out.append(struct.pack("<Q", e.shape[0]))
        for d in range(e.shape[0]):
            out.append(save_ciphertext(pid, e[d], N, k, 2, True))
    return b"".join(out)


def load_kswitch(f):
    pid = f.read(32)
    dbc, dim1 = struct.unpack("<iQ", f.read(12))
    entries = []
    for _ in range(dim1):
        (dim2,) = struct.unpack("<Q", f.read(8))
        if dim2 == 0:
            entries.append(None)
            continue
        entries.append(np.stack([load_ciphertext(f)["words"] for _ in range(dim2)]))
    return dict(parms_id=pid, dbc=dbc, entries=entries)


def save_environment(N, q, t, dbc_relin, dbc_galois, pk, rlk, glk, sk):
    """pk [2kN], rlk [D][2kN], glk {elt: [D][2kN]}, sk [kN] or None -- all NTT form (the library's key export layout)"""
    k = len(q)
    pid = parms_id(N, q, t)
    gal = [None] * N
    for elt, key in glk.items():
        gal[(int(elt) - 1) >> 1] = key
    parts = [save_parms(N, q, t), save_ciphertext(pid, pk, N, k, 2, True), save_kswitch(pid, dbc_relin, [rlk], N, k), save_kswitch(pid, dbc_galois, gal, N, k)]
    parts.append(save_plaintext(sk, pid) if sk is not None else save_plaintext(np.zeros(0, np.uint64)))
    return b"".join(parts)


def load_environment(raw):
    f = io.BytesIO(raw)
    N, q, t, sd = load_parms(f)
    pk = load_ciphertext(f)
    rlk = load_kswitch(f)
    glk = load_kswitch(f)
    sk = load_plaintext(f)
    assert f.read() == b""
    return dict(N=N, q=q, t=t, pk=pk["words"], dbc_relin=rlk["dbc"], rlk=rlk["entries"][0], dbc_galois=glk["dbc"],
                glk={2 * i + 1: e for i, e in enumerate(glk["entries"]) if e is not None}, sk=sk["words"] if sk["words"].size else None,
                parms_id=pk["parms_id"])


def save_archive(envs):
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_STORED) as z:
        for i, e in enumerate(envs):
            z.writestr("environment%03d" % i, e)
    return buf.getvalue()


def load_archive(raw):
    with zipfile.ZipFile(io.BytesIO(raw)) as z:
        names = sorted(n for n in z.namelist() if n.startswith("environment"))
        return [z.read(n) for n in names]


def fmt_double(x):
    """.NET Framework Double.ToString(): 15 significant digits, shortest form"""
    return "%.15g" % x


def write_atomic(scale, signed, fmt_name, dim, encrypted, blobs, nl="\r\n"):
    lines = ["<Start EncryptedVector>", fmt_double(scale), "True" if signed else "False", fmt_name, str(dim), "Encrypted" if encrypted else "Plain",
             str(len(blobs)), base64.b64encode(b"".join(blobs)).decode(), "<End EncryptedVector>"]
    return nl.join(lines) + nl


def write_vector(scale, atomics, nl="\r\n"):
    return "<Start LargeEncryptedVector>" + nl + fmt_double(scale) + nl + str(len(atomics)) + nl + "".join(atomics) + "<End LargeEncryptedVector>" + nl


def read_vector(text):
    """-> dict(scale, channels=[dict(scale, signed, format, dim, encrypted, count, blob)])"""
    lines = text.replace("\r\n", "\n").split("\n")
    it = iter(lines)
    assert next(it) == "<Start LargeEncryptedVector>"
    scale = float(next(it))
    n = int(next(it))
    ch = []
    for _ in range(n):
        assert next(it) == "<Start EncryptedVector>"
        d = dict(scale=float(next(it)), signed=next(it) == "True", format=next(it), dim=int(next(it)), encrypted=next(it) == "Encrypted", count=int(next(it)))
        d["blob"] = base64.b64decode(next(it))
        assert next(it) == "<End EncryptedVector>"
        ch.append(d)
    assert next(it) == "<End LargeEncryptedVector>"
    return dict(scale=scale, channels=ch)
