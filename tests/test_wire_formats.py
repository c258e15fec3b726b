"""Wire / on-disk formats: the product's C++ implementation (csrc/wire.cu, through the C ABI) against the independent
Python restatement in oracle/wire_py.py, plus the reference's two serialisation tests restated:
`SaveLoadKeys` and `SaveAndLoadMatrix` ("HE Wrapper Tests/BasicOperations.cs:291-331").  The SEAL 3.2 binary layouts themselves are
unpinned against the real SEAL binary (not in the reference tree; see the header of csrc/wire.cu)."""
import io
import zipfile

import numpy as np
import pytest

from oracle import wire_py as W


def test_python_restatement_round_trips():
    N, q, t = 4096, [0xffffee001, 0xffffc4001, 0x1ffffe0001], 40961
    k = len(q)
    rng = np.random.default_rng(0)
    pk = rng.integers(0, 1 << 36, 2 * k * N, dtype=np.uint64)
    rlk = rng.integers(0, 1 << 36, (12, 2 * k * N), dtype=np.uint64)
    glk = {3: rng.integers(0, 1 << 36, (6, 2 * k * N), dtype=np.uint64), 2 * N - 1: rng.integers(0, 1 << 36, (6, 2 * k * N), dtype=np.uint64)}
    sk = rng.integers(0, 1 << 36, k * N, dtype=np.uint64)
    for with_sk in (True, False):
        env = W.save_environment(N, q, t, 10, 20, pk, rlk, glk, sk if with_sk else None)
        back = W.load_environment(W.load_archive(W.save_archive([env]))[0])
        assert back["N"] == N and back["q"] == q and back["t"] == t and back["dbc_relin"] == 10 and back["dbc_galois"] == 20
        assert np.array_equal(back["pk"], pk) and np.array_equal(back["rlk"], rlk)
        assert sorted(back["glk"]) == sorted(glk) and all(np.array_equal(back["glk"][e], glk[e]) for e in glk)
        assert (back["sk"] is None) == (not with_sk) and (not with_sk or np.array_equal(back["sk"], sk))
        assert back["parms_id"] == W.parms_id(N, q, t) and len(back["parms_id"]) == 32
    assert W.fmt_double(12.0) == "12" and W.fmt_double(0.0625) == "0.0625" and W.fmt_double(1e20) == "1e+20"


@pytest.fixture(scope="module")
def factory():
    from cryptonets_b200.he import B200BfvFactory
    f = B200BfvFactory([40961, 65537], 4096, seed=11)
    yield f
    f.Dispose()


@pytest.mark.gpu
def test_key_archive_matches_the_python_restatement(factory):
    eng = factory.engine
    raw = eng.save_keys(with_private_keys=True)
    entries = W.load_archive(raw)
    with zipfile.ZipFile(io.BytesIO(raw)) as z:  # a stock ZIP reader accepts the container; entries are stored, named like the reference's
        assert z.namelist() == ["environment000", "environment001"] and z.testzip() is None
        assert all(i.compress_type == zipfile.ZIP_STORED for i in z.infolist())
    for ch, blob in enumerate(entries):
        env = W.load_environment(blob)
        assert (env["N"], env["q"], env["t"]) == (eng.N, eng.q, eng.primes[ch])
        assert (env["dbc_relin"], env["dbc_galois"]) == (10, 20)
        assert env["parms_id"] == W.parms_id(eng.N, eng.q, eng.primes[ch])
        assert np.array_equal(env["pk"], eng.export_key(ch, 1)) and np.array_equal(env["sk"], eng.export_key(ch, 0))
        assert np.array_equal(env["rlk"].ravel(), eng.export_key(ch, 2))
        assert set(env["glk"]) == set(eng.galois_elts())  # (the list names 3^(N/4) twice: it is its own inverse)
        for elt in eng.galois_elts()[:3]:
            assert np.array_equal(env["glk"][elt].ravel(), eng.export_key(ch, 3, elt))
        # byte for byte: the Python writer fed with the exported keys reproduces the entry
        glk = {e: eng.export_key(ch, 3, e) for e in set(eng.galois_elts())}
        again = W.save_environment(eng.N, eng.q, eng.primes[ch], 10, 20, eng.export_key(ch, 1), eng.export_key(ch, 2).reshape(eng.relin_digits, -1),
                                   {e: g.reshape(eng.galois_digits, -1) for e, g in glk.items()}, eng.export_key(ch, 0))
        assert again == blob
    pub = W.load_environment(W.load_archive(eng.save_keys(with_private_keys=False))[0])
    assert pub["sk"] is None


@pytest.mark.gpu
def test_save_load_keys(factory, tmp_path):
    """BasicOperations.SaveLoadKeys: save with private keys, encrypt with the first factory, decrypt with a factory built from the file."""
    from cryptonets_b200.he import B200BfvFactory
    from cryptonets_b200.interfaces import EVectorFormat
    path = str(tmp_path / "keys2.keys")
    factory.Save(str(tmp_path / "keys.keys"))
    factory.Save(path, True)
    v = np.array([1.0, 2.0, 3.0])
    vEnc = factory.GetEncryptedVector(v, EVectorFormat.dense, 1)
    factory2 = B200BfvFactory(path)
    try:
        assert factory2.engine.primes == factory.engine.primes and factory2.engine.q == factory.engine.q
        buf = io.StringIO()
        vEnc.Write(buf)  # the ciphertexts travel as the reference's text form (a vector here is a device handle of one context)
        buf.seek(0)
        w = factory2.LoadVector(buf).Decrypt()
        assert np.array_equal(np.asarray(w), v)
        # the loaded factory is fully functional: encrypt + evaluate + decrypt
        a = factory2.GetEncryptedVector(np.array([4.0, -5.0, 6.0]), EVectorFormat.dense, 1)
        assert np.array_equal(np.asarray(a.PointwiseMultiply(a).Decrypt()), [16.0, 25.0, 36.0])
    finally:
        factory2.Dispose()
    # an archive written without the secret key (deflate framing with stored blocks, as .NET's NoCompression emits): encrypts, cannot decrypt
    entries = W.load_archive(factory.engine.save_keys(False))
    zbuf = io.BytesIO()
    with zipfile.ZipFile(zbuf, "w", zipfile.ZIP_DEFLATED, compresslevel=0) as z:
        for i, e in enumerate(entries):
            z.writestr("environment%03d" % i, e)
    factory3 = B200BfvFactory(zbuf.getvalue())
    try:
        c = factory3.GetEncryptedVector(v, EVectorFormat.dense, 1)
        with pytest.raises(Exception, match="secret key"):
            c.Decrypt()
        buf = io.StringIO()
        c.Write(buf)
        buf.seek(0)
        assert np.array_equal(np.asarray(factory.LoadVector(buf).Decrypt()), v)  # the key owner decrypts what the public-key holder encrypted
    finally:
        factory3.Dispose()


@pytest.mark.gpu
def test_save_and_load_matrix(factory):
    """BasicOperations.SaveAndLoadMatrix: Write -> LoadMatrix -> equal decryptions; encrypted and plain, dense and sparse."""
    from cryptonets_b200.interfaces import EMatrixFormat, EVectorFormat
    rng = np.random.default_rng(3)
    m = rng.integers(-100, 100, (5, 7)).astype(np.float64)
    mat = factory.GetEncryptedMatrix(m, EMatrixFormat.ColumnMajor, 12)
    buf = io.StringIO()
    mat.Write(buf)
    text = buf.getvalue()
    assert text.startswith("<Start LargeEncryptedMatrix>\r\nColumnMajor\r\n7\r\n<Start LargeEncryptedVector>\r\n12\r\n2\r\n<Start EncryptedVector>\r\n1\r\nFalse\r\ndense\r\n5\r\nEncrypted\r\n1\r\n")
    buf.seek(0)
    mat2 = factory.LoadMatrix(buf)
    assert mat2.Format == EMatrixFormat.ColumnMajor and np.array_equal(np.asarray(mat.Decrypt()), np.asarray(mat2.Decrypt()))
    # the text parses with the independent reader and carries the exact ciphertext words
    eng = factory.engine
    first = text.split("<Start LargeEncryptedVector>")[1]
    parsed = W.read_vector("<Start LargeEncryptedVector>" + first.split("<End LargeEncryptedVector>")[0] + "<End LargeEncryptedVector>\r\n")
    assert parsed["scale"] == 12 and len(parsed["channels"]) == 2
    for ch, d in enumerate(parsed["channels"]):
        ct = W.load_ciphertext(io.BytesIO(d["blob"]))
        assert ct["parms_id"] == W.parms_id(eng.N, eng.q, eng.primes[ch]) and not ct["ntt"] and ct["size"] == 2
        assert np.array_equal(ct["words"], mat.vectors[0].vec.export_raw(ch, 0))
    for fmt in (EVectorFormat.dense, EVectorFormat.sparse):  # plain vectors: Plaintext streams
        p = factory.GetPlainVector(np.array([3.0, -4.0, 5.0]), fmt, 2)
        b2 = io.StringIO()
        p.Write(b2)
        b2.seek(0)
        q = factory.LoadVector(b2)
        assert not q.IsEncrypted and q.Format == fmt and q.Scale == 2 and np.array_equal(np.asarray(q.Decrypt()), [3.0, -4.0, 5.0])
    with pytest.raises(Exception, match="Bad stream format"):
        factory.LoadVector(io.StringIO("<Start SomethingElse>\n"))
    mat.Dispose()
    mat2.Dispose()
