"""B200 backend of the CryptoNets plugin API: IFactory / IVector / IMatrix / IComputationEnvironment.

Python mirror of the C# shim a maintainer would add next to `EncryptedSealBfvFactory`: same member names,
argument meaning and exception behaviour as `HE Wrapper/IFactory.cs:20-130`, `IVector.cs:20-136`, `IMatrix.cs:18-122`, with
every method a call into libcnhe.so.  Vectors hold device handles instead of SEAL `Ciphertext[]`.
Errors surface as Python exceptions carrying the reference's message (the reference throws System.Exception)."""
import numpy as np

from .engine import ALL_SLOTS, Engine, Vec
from .interfaces import EMatrixFormat, EVectorFormat


class B200BfvEnvironment:
    """IComputationEnvironment (`HE Wrapper/IComputationEnvironment.cs:12-24`).  All environments of a factory share the
    context's CUDA stream; the object exists so that reference call sites keep their shape."""

    def __init__(self, factory):
        self.ParentFactory = factory

    Primes = property(lambda s: list(s.ParentFactory.engine.primes))


class B200BfvVector:
    """IVector over a device-resident cnhe_vec (== EncryptedSealBfvVector, `EncryptedSealBfvVector.cs:150-573`)."""

    def __init__(self, factory, vec):
        self.factory = factory
        self.vec = vec
        self.IsSigned = True

    eng = property(lambda s: s.factory.engine)
    Dim = property(lambda s: s.vec.dim)
    Scale = property(lambda s: s.vec.scale)
    Format = property(lambda s: EVectorFormat(s.vec.format))
    IsEncrypted = property(lambda s: s.vec.is_encrypted)
    BlockSize = property(lambda s: s.eng.N)
    Data = property(lambda s: s.vec)

    def _wrap(self, vec):
        return B200BfvVector(self.factory, vec)

    def Dispose(self):
        if self.vec is not None:
            self.vec.dispose()
            self.vec = None

    def Write(self, writer):
        """IVector.Write(StreamWriter) (`EncryptedSealBfvVector.cs:428-437`)."""
        writer.write(self.eng.write_vector(self.vec))

    def RegisterScale(self, scale):
        self.vec.register_scale(scale)

    def RegisterDim(self, dim):
        self.vec.register_dim(dim)

    def Decrypt(self, env=None):
        return self.eng.decrypt(self.vec)

    def DecryptFullPrecision(self, env=None):
        """IVector.DecryptFullPrecision (`EncryptedSealBfvVector.cs:343-348,397-411`): exact integers (Python ints stand in for BigInteger),
        the CRT join of the per-modulus residues, centred when the vector is signed; NOT divided by Scale (the reference does not)."""
        f, res = self.factory, self.eng.decrypt_residues(self.vec)
        out = []
        for j in range(res.shape[1]):
            x = sum(int(c) * int(res[i, j]) for i, c in enumerate(f.preComputedCoefficients)) % f.bigFactor
            if self.IsSigned and x * 2 > f.bigFactor:
                x -= f.bigFactor
            out.append(x)
        return out

    def Add(self, v, env=None):
        return self._wrap(self.eng.add(self.vec, v.vec))

    def Subtract(self, v, env=None):
        return self._wrap(self.eng.sub(self.vec, v.vec))

    def PointwiseMultiply(self, v, env=None):
        return self._wrap(self.eng.pointwise_multiply(self.vec, v.vec))

    def DotProduct(self, v, env=None, length=None, ForceOutputInColumn=None):
        fc = -1 if ForceOutputInColumn is None else ForceOutputInColumn
        return self._wrap(self.eng.dot_product(self.vec, v.vec, ALL_SLOTS if length is None else length, fc))

    def SumAllSlots(self, env=None, length=None, ForceOutputInColumn=None):
        fc = -1 if ForceOutputInColumn is None else ForceOutputInColumn
        return self._wrap(self.eng.sum_all_slots(self.vec, ALL_SLOTS if length is None else length, fc))

    def Duplicate(self, count, env=None):
        return self._wrap(self.eng.duplicate(self.vec, count))

    def Rotate(self, amount, env=None):
        return self._wrap(self.eng.rotate(self.vec, amount))

    def Permute(self, selections, shifts, outputDim, env=None):
        sel = [None if s is None else s.vec for s in selections]
        return self._wrap(self.eng.permute(self.vec, sel, shifts, outputDim))


class B200BfvMatrix:
    """IMatrix as an array of vectors (`HE Wrapper/EncryptedSealBfvMatrix.cs:14-231`)."""

    def __init__(self, factory, vectors, fmt=EMatrixFormat.ColumnMajor, CopyVectors=True):
        if vectors and any(v.Dim != vectors[0].Dim for v in vectors):
            raise Exception("all columns of a matrix should have the same size")
        self.factory = factory
        self.vectors = [factory.CopyVector(v) for v in vectors] if CopyVectors else list(vectors)
        self.Format = fmt
        self.DataDisposedExternaly = False
        self.Batched = True  # False replays the reference's per-row call sequence

    eng = property(lambda s: s.factory.engine)
    RowCount = property(lambda s: len(s.vectors) if s.Format == EMatrixFormat.RowMajor else s.vectors[0].Dim)
    ColumnCount = property(lambda s: len(s.vectors) if s.Format == EMatrixFormat.ColumnMajor else s.vectors[0].Dim)
    Scale = property(lambda s: s.vectors[0].Scale)
    BlockSize = property(lambda s: s.eng.N)
    IsEncrypted = property(lambda s: all(v.IsEncrypted for v in s.vectors))
    Data = property(lambda s: s.vectors)

    def Write(self, writer):
        """IMatrix.Write(StreamWriter) (`EncryptedSealBfvMatrix.cs:199-208`)."""
        nl = "\r\n"
        writer.write("<Start LargeEncryptedMatrix>" + nl + EMatrixFormat(self.Format).name + nl + str(len(self.vectors)) + nl)
        for v in self.vectors:
            v.Write(writer)
        writer.write("<End LargeEncryptedMatrix>" + nl)

    def Dispose(self):
        if self.vectors is not None and not self.DataDisposedExternaly:
            vs = [v for v in self.vectors if v is not None and v.vec is not None]
            if vs:
                self.eng.dispose_many([v.vec for v in vs])
            for v in vs:
                v.vec = None
        self.vectors = None

    def RegisterScale(self, scale):
        for v in self.vectors:
            v.RegisterScale(scale)

    def Decrypt(self, env=None):
        rows = self.eng.decrypt_many([v.vec for v in self.vectors])
        return rows if self.Format == EMatrixFormat.RowMajor else rows.T

    def MulRows(self, v, ForceDenseFormat, first_row, total_rows):
        """Row-major product for a slice of the rows held by this matrix (global rows first_row..): the per-rank piece of a row-sharded
        dense layer (cnhe_mat_mul_rowmajor_shard; cryptonets_b200/parallel.py combines the pieces)."""
        if self.Format != EMatrixFormat.RowMajor:
            raise Exception("MulRows expects a RowMajor matrix")
        return B200BfvVector(self.factory, self.eng.mat_mul_rowmajor_shard([r.vec for r in self.vectors], v.vec, ForceDenseFormat, first_row, total_rows))

    def Mul(self, v, env=None, ForceDenseFormat=False):
        f = self.factory
        if self.Format == EMatrixFormat.ColumnMajor:
            if ForceDenseFormat:
                raise Exception("Forcing dense format is available only in RowMajor mode")
            return B200BfvVector(f, self.eng.mat_mul_colmajor_sparse([c.vec for c in self.vectors], v.vec))
        if self.Batched and v.IsEncrypted and not self.IsEncrypted and v.vec.blocks == 1:
            # all rows through each stage together (same ciphertexts as the per-row loop below)
            return B200BfvVector(f, self.eng.mat_mul_rowmajor([r.vec for r in self.vectors], v.vec, ForceDenseFormat))
        if not ForceDenseFormat:  # EncryptedSealBfvMatrix.cs:79-89
            tmp = [row.DotProduct(v, env) for row in self.vectors]
            res = B200BfvVector(f, self.eng.generate_sparse_of_array([t.vec for t in tmp]))
            for t in tmp:
                t.Dispose()
            return res
        total = None  # EncryptedSealBfvMatrix.cs:90-120
        for i, row in enumerate(self.vectors):
            t = row.DotProduct(v, env, ForceOutputInColumn=i)
            if total is None:
                total = t
            else:
                s = total.Add(t, env)
                total.Dispose()
                t.Dispose()
                total = s
        total.RegisterDim(len(self.vectors))
        if total.Format != EVectorFormat.dense:
            raise Exception("Internal probloem: expecting the output to be dense")
        return total

    def _check(self, m):
        if m.Format != self.Format:
            raise Exception("Format mismatch")
        if m.RowCount != self.RowCount:
            raise Exception("Row count mismatch")
        if m.ColumnCount != self.ColumnCount:
            raise Exception("Column count mismatch")

    def Add(self, m, env=None):
        self._check(m)
        out = [a.Add(b, env) for a, b in zip(self.vectors, m.vectors)]
        return B200BfvMatrix(self.factory, out, self.Format, CopyVectors=False)

    def ElementWiseMultiply(self, m, env=None):
        self._check(m)
        if m is self:  # SquareActivation: one batched wave over every column
            out = self.eng.layer_square([v.vec for v in self.vectors])
            return B200BfvMatrix(self.factory, [B200BfvVector(self.factory, o) for o in out], self.Format, CopyVectors=False)
        out = [a.PointwiseMultiply(b, env) for a, b in zip(self.vectors, m.vectors)]
        return B200BfvMatrix(self.factory, out, self.Format, CopyVectors=False)

    def GetColumn(self, i):
        if i >= len(self.vectors):
            raise Exception("Column does not exist")
        if self.Format != EMatrixFormat.ColumnMajor:
            raise Exception("Columns can be extracted only from a column major matrix")
        return self.vectors[i]

    def GetRow(self, i):
        if i >= len(self.vectors):
            raise Exception("Row does not exist")
        if self.Format != EMatrixFormat.RowMajor:
            raise Exception("Rows can be extracted only from a row major matrix")
        return self.vectors[i]

    def SetColumn(self, i, vector):
        if i >= len(self.vectors):
            raise Exception("Column does not exist")
        if self.Format != EMatrixFormat.ColumnMajor:
            raise Exception("Columns can be set only from a column major matrix")
        if vector.Dim != self.vectors[i].Dim:
            raise Exception("dimension of vector does not match the dimension of the vector it is replacing")
        if vector.Scale != self.vectors[i].Scale:
            raise Exception("Scale of vector does not match the scale of the vector it is replacing")
        if vector.IsEncrypted != self.vectors[i].IsEncrypted:
            raise Exception("can't exchange encrypted and not encrypted vectors")
        self.vectors[i] = vector

    def ConvertToColumnVector(self, env=None):
        return B200BfvVector(self.factory, self.eng.stack([v.vec for v in self.vectors]))

    def Interleave(self, shift, env=None):
        if self.Format != EMatrixFormat.ColumnMajor:
            raise Exception("Expecting ColumnMajor matrix")
        return B200BfvVector(self.factory, self.eng.interleave([v.vec for v in self.vectors], shift))


def _read_block(reader, end_marker):
    """lines of a text stream up to and including the line `end_marker`"""
    out = []
    while True:
        line = reader.readline()
        if not line:
            raise Exception("Bad stream format.")
        out.append(line)
        if line.rstrip("\r\n") == end_marker:
            return "".join(out)


class B200BfvFactory:
    """IFactory (`HE Wrapper/IFactory.cs:20-130`); constructor arguments of EncryptedSealBfvFactory (`:247-260`)."""

    DefaultDecompositionBitCount = 10
    DefaultGaloisDecompositionBitCount = 20

    def __init__(self, primes=None, n=4096, DecompositionBitCount=10, GaloisDecompositionBitCount=20, SmallModulusCount=-1, seed=None,
                 device=0, generate_keys=True):
        """seed=None (default): keys and encryption randomness from the OS CSPRNG, as SEAL's KeyGenerator/Encryptor give the reference.
        An integer seed selects the deterministic sampler shared with the CPU oracle: parity tests only."""
        if isinstance(primes, (str, bytes, bytearray)):  # EncryptedSealBfvFactory(fileName) (IFactory.cs:262-265): parameters and keys from a key archive
            data = open(primes, "rb").read() if isinstance(primes, str) else bytes(primes)
            self.engine = Engine(None, archive=data, device=device)
            primes = self.engine.primes
        else:
            if primes is None:
                primes, n = [40961, 65537, 114689, 147457, 188417], 4096  # IFactory.cs:247-253
            self.engine = Engine(primes, n, DecompositionBitCount, GaloisDecompositionBitCount, SmallModulusCount, device)
            if generate_keys:
                self.engine.keygen(seed)
        self._env = B200BfvEnvironment(self)
        big = 1
        for p in primes:
            big *= int(p)
        self.bigFactor = big
        self.preComputedCoefficients = [(big // int(p)) * pow((big // int(p)) % int(p), -1, int(p)) for p in primes]

    Primes = property(lambda s: list(s.engine.primes))

    def AllocateComputationEnv(self):
        return self._env

    def FreeComputationEnv(self, env):
        pass

    def _big(self, v, fmt, encrypt):  # the IEnumerable<BigInteger> overloads (IFactory.cs:29,43): SplitBigNumbers on exact integers
        vals = [int(x) % self.bigFactor for x in v]
        res = np.array([[x % int(p) for x in vals] for p in self.engine.primes], dtype=np.uint64)
        return B200BfvVector(self, self.engine.from_residues(res, 1.0, int(fmt), encrypt))

    def GetPlainVector(self, v, fmt, scale=None):
        if scale is None:
            return self._big(v, fmt, False)
        return B200BfvVector(self, self.engine.plain(np.asarray(v, dtype=np.float64), scale, int(fmt)))

    def GetEncryptedVector(self, v, fmt, scale=None):
        if scale is None:
            return self._big(v, fmt, True)
        return B200BfvVector(self, self.engine.encrypt(np.asarray(v, dtype=np.float64), scale, int(fmt)))

    def CopyVector(self, v):
        return B200BfvVector(self, self.engine.copy(v.vec))

    def _rows(self, m, fmt):
        m = np.asarray(m, dtype=np.float64)
        return m.T if fmt == EMatrixFormat.ColumnMajor else m

    def GetPlainMatrix(self, m, fmt, scale):
        vecs = [self.GetPlainVector(r, EVectorFormat.dense, scale) for r in self._rows(m, fmt)]
        return B200BfvMatrix(self, vecs, fmt, CopyVectors=False)

    def GetEncryptedMatrix(self, m, fmt, scale):
        rows = np.ascontiguousarray(self._rows(m, fmt))
        vecs = [B200BfvVector(self, v) for v in self.engine.encrypt_many(rows, scale)]
        return B200BfvMatrix(self, vecs, fmt, CopyVectors=False)

    def GetMatrix(self, vectors, fmt, CopyVectors=True):
        return B200BfvMatrix(self, vectors, fmt, CopyVectors=CopyVectors)

    # ---- wire formats (IFactory.cs:474-495; SEAL streams unpinned, see csrc/wire.cu)
    def Save(self, target, withPrivateKeys=False):
        """IFactory.Save(stream | fileName, withPrivateKeys): the ZIP key archive."""
        data = self.engine.save_keys(withPrivateKeys)
        if isinstance(target, str):
            with open(target, "wb") as f:
                f.write(data)
        else:
            target.write(data)
        return target

    def LoadVector(self, reader):
        """IFactory.LoadVector(StreamReader): `reader` is a text stream positioned at "<Start LargeEncryptedVector>"."""
        text = _read_block(reader, "<End LargeEncryptedVector>")
        vec, _ = self.engine.read_vector(text)
        return B200BfvVector(self, vec)

    def LoadMatrix(self, reader):  # EncryptedSealBfvMatrix.Read (EncryptedSealBfvMatrix.cs:182-197)
        if reader.readline().rstrip("\r\n") != "<Start LargeEncryptedMatrix>":
            raise Exception("Bad stream format.")
        fmt = EMatrixFormat[reader.readline().strip()]
        n = int(reader.readline())
        vecs = [self.LoadVector(reader) for _ in range(n)]
        if reader.readline().rstrip("\r\n") != "<End LargeEncryptedMatrix>":
            raise Exception("Bad stream format.")
        return B200BfvMatrix(self, vecs, fmt, CopyVectors=False)

    def GetValueFromString(self, s):  # IFactory.cs:395-403
        f = [int(x) for x in s.split(",")]
        return sum(c * x for c, x in zip(self.preComputedCoefficients, f)) % self.bigFactor

    def GetStringFromValue(self, value):
        return ",".join(str(int(value) % int(p)) for p in self.engine.primes)

    # fused layer entry point used by PoolLayer.Apply (NeuralNetworks/PoolLayer.cs:149-229)
    def ConvDenseLayer(self, inputs, gather, weights, bias, M, K):
        out = self.engine.layer_conv_dense([v.vec for v in inputs], gather, [w.vec for w in weights],
                                           None if bias is None else [b.vec for b in bias], M, K)
        return [B200BfvVector(self, o) for o in out]

    def Dispose(self):
        self.engine.close()
