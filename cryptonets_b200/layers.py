"""Layer API of the reference (`NeuralNetworks/`), restated over the IFactory / IMatrix / IVector interfaces.

Same class and property names as the C# layer library so that the network definitions of `CryptoNets/CryptoNets.cs:20-75`
and `LowLatencyCryptoNets/LoLaCryptonets.cs:280-329` read the same.  Layers talk only to the plugin interfaces; with a
B200BfvFactory, PoolLayer.Apply issues one fused device call for the whole layer instead of the reference's per-output
fan-out (`PoolLayer.cs:196-227`), with identical outputs; pass Fused=False to replay the reference's call sequence."""
import time

import numpy as np

from .interfaces import EMatrixFormat, EVectorFormat
from .raw import RawFactory, RawMatrix


class ConvolutionEngine:
    """Index arithmetic of `NeuralNetworks/ConvolutionEngine.cs:10-145`."""

    def __init__(self):
        self.InputShape = None
        self._kernel = None
        self.Stride = None
        self.Padding = None
        self.Upperpadding = None
        self.Lowerpadding = None
        self.MapCount = None
        self.Offsets = None
        self.Corners = None
        self.prepared = False

    @property
    def KernelShape(self):
        return self._kernel

    @KernelShape.setter
    def KernelShape(self, value):
        self._kernel = list(value)
        offs, off = [], [0] * len(value)
        while True:  # first axis fastest (:39-54)
            offs.append(list(off))
            go = False
            for i in range(len(value)):
                off[i] += 1
                if off[i] < value[i]:
                    go = True
                    break
                off[i] = 0
            if not go:
                break
        self.Offsets = offs

    def Prepare(self):
        if self.prepared:
            return
        n = len(self.InputShape)
        self.Upperpadding = self.Upperpadding or [0] * n
        self.Lowerpadding = self.Lowerpadding or [0] * n
        self.Padding = self.Padding or [False] * n
        self.maps = int(np.prod(self.MapCount)) if self.MapCount is not None else 1
        ks = self._kernel
        lo = [-self.Lowerpadding[i] - (-(ks[i] // 2) if self.Padding[i] else 0) for i in range(n)]
        hi = [self.InputShape[i] + self.Upperpadding[i] - (((ks[i] + 1) // 2) if self.Padding[i] else ks[i]) for i in range(n)]
        corners, cur = [], list(lo)
        while True:  # last axis fastest (:61-79)
            corners.append(list(cur))
            go = False
            for i in range(n - 1, -1, -1):
                cur[i] += self.Stride[i]
                if cur[i] <= hi[i]:
                    go = True
                    break
                cur[i] = lo[i]
            if not go:
                break
        self.Corners = corners
        self.prepared = True

    def Location(self, corner, offset, shape, bias=0):
        if not self.prepared:
            self.Prepare()
        index = 0
        for i in range(len(offset)):
            cord = (corner[i] + offset[i]) if corner is not None else offset[i]
            if cord < 0 or cord >= shape[i]:
                return -1
            index = index * shape[i] + cord
        return index + bias

    def GetDenseWeights(self, weights):  # :121-144
        """The convolution as a dense [maps*corners x prod(InputShape)] row-major matrix (CIFAR: 5488 x 16268), vectorised."""
        self.Prepare()
        weights = np.asarray(weights, dtype=np.float64)
        shape = np.asarray(self.InputShape)
        nc, cols = len(self.Corners), int(np.prod(self.InputShape))
        ksize = int(np.prod(self._kernel))
        offs = np.asarray(self.Offsets)                                      # [O, n]
        coords = np.asarray(self.Corners)[:, None, :] + offs[None, :, :]      # [C, O, n]
        valid = np.all((coords >= 0) & (coords < shape), axis=2)
        loc = np.zeros(coords.shape[:2], dtype=np.int64)
        kidx = np.zeros(len(offs), dtype=np.int64)
        for i in range(len(shape)):                                          # same row-major index as Location()
            loc = loc * shape[i] + coords[:, :, i]
            kidx = kidx * self._kernel[i] + offs[:, i]
        ci, oi = np.nonzero(valid)
        mat = np.zeros((self.maps * nc, cols))
        for m in range(self.maps):
            mat[m * nc + ci, loc[ci, oi]] = weights[kidx[oi] + m * ksize]
        return mat.reshape(-1)

    def GetDenseBias(self, bias):
        return np.repeat(np.asarray(bias, dtype=np.float64)[: self.maps], len(self.Corners))


class BaseLayer:
    """`NeuralNetworks/BaseLayer.cs:9-105` (pull model: GetNext() asks the Source, then applies this layer)."""

    def __init__(self, **kw):
        self.Source = None
        self._factory = None
        self.Verbose = False
        self.layerPrepared = False
        self.LastSeconds = None
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def Factory(self):
        return self._factory if self._factory is not None else self.Source.Factory

    @Factory.setter
    def Factory(self, f):
        self._factory = f

    def Apply(self, m):
        raise NotImplementedError

    def GetNext(self):
        if not self.layerPrepared:
            self.Prepare()
            self.layerPrepared = True
        m = self.Source.GetNext()
        start = time.time()
        res = self.Apply(m)
        self.LastSeconds = time.time() - start
        if self.Verbose:
            print("Layer %s computed in %.4f seconds layer width (%d,%d)" % (type(self).__name__, self.LastSeconds, m.RowCount, m.ColumnCount))
        if res is not m:
            m.Dispose()
        return res

    def GetOutputScale(self):
        return self.Source.GetOutputScale()

    def OutputDimension(self):
        return self.Source.OutputDimension()

    def Prepare(self):
        pass

    def PrepareNetwork(self):
        if self.Source is not None:
            self.Source.PrepareNetwork()
        self.Prepare()
        self.layerPrepared = True

    def DisposeNetwork(self):
        if self.Source is not None:
            self.Source.DisposeNetwork()
        self.Dispose()

    def Dispose(self):
        pass


class MatrixSource(BaseLayer):
    """Input layer over an in-memory feature matrix: the role of `BatchReader` (`BatchReader.cs:59-109`) with the TSV parsing
    replaced by a synthetic / caller-supplied array (no dataset ships with the reference).  rows = images."""

    def __init__(self, features, Scale=1.0, NormalizationFactor=1.0, MaxSlots=None, labels=None):
        super().__init__()
        self.features = np.asarray(features, dtype=np.float64)
        self.Scale = Scale
        self.NormalizationFactor = NormalizationFactor
        self.MaxSlots = MaxSlots or len(self.features)
        self.Labels = labels
        self._factory = RawFactory(8192)
        self.pos = 0

    def PrepareNetwork(self):
        pass

    def DisposeNetwork(self):
        pass

    def GetNext(self):
        if self.pos >= len(self.features):
            self.pos = 0
        chunk = self.features[self.pos: self.pos + self.MaxSlots]
        self.pos += self.MaxSlots
        return RawMatrix(chunk * self.NormalizationFactor, self.Scale, EMatrixFormat.ColumnMajor, 0)

    def GetOutputScale(self):
        return self.Scale

    def OutputDimension(self):
        return self.features.shape[1]


class LLConvReader(MatrixSource):
    """One image -> im2col matrix [corners x offsets] (`LLConvReader.cs:96-158`)."""

    def __init__(self, features, Scale, NormalizationFactor, InputShape, KernelShape, Stride, Upperpadding=None, Lowerpadding=None, Padding=None):
        super().__init__(features, Scale, NormalizationFactor, MaxSlots=1)
        self.ce = ConvolutionEngine()
        self.ce.InputShape, self.ce.KernelShape, self.ce.Stride = InputShape, KernelShape, Stride
        self.ce.Upperpadding, self.ce.Lowerpadding, self.ce.Padding = Upperpadding, Lowerpadding, Padding
        self.ce.Prepare()

    def GetNext(self):
        if self.pos >= len(self.features):
            self.pos = 0
        f = self.features[self.pos] * self.NormalizationFactor
        self.pos += 1
        ce = self.ce
        mat = np.zeros((len(ce.Corners), len(ce.Offsets)))
        for c, corner in enumerate(ce.Corners):
            for o, off in enumerate(ce.Offsets):
                l = ce.Location(corner, off, ce.InputShape)
                mat[c, o] = f[l] if l >= 0 else 0
        return RawMatrix(mat, self.Scale, EMatrixFormat.ColumnMajor, 0)

    def OutputDimension(self):
        return len(self.ce.Corners)


class EncryptLayer(BaseLayer):
    """`NeuralNetworks/EncryptLayer.cs:10-20`"""

    def Apply(self, m):
        res = self.Factory.GetEncryptedMatrix(m.Data, EMatrixFormat.ColumnMajor, 1)
        res.RegisterScale(m.Scale)
        return res


class TimingLayer(BaseLayer):
    """`NeuralNetworks/TimingLayer.cs:15-66`; device work is asynchronous, so a counter boundary synchronises the factory first."""

    TotalTimeMS, N, StartTime = {}, {}, {}

    def __init__(self, StartCounters=(), StopCounters=(), **kw):
        super().__init__(**kw)
        self.StartCounters, self.StopCounters = list(StartCounters), list(StopCounters)

    @classmethod
    def Reset(cls):
        cls.TotalTimeMS.clear(); cls.N.clear(); cls.StartTime.clear()

    @classmethod
    def GetStats(cls):
        return "\t".join("%s %.2f" % (k, v / cls.N[k]) for k, v in cls.TotalTimeMS.items())

    def Apply(self, m):
        eng = getattr(self.Factory, "engine", None)
        if eng is not None:
            eng.sync()
        now = time.time()
        for c in self.StartCounters:
            TimingLayer.StartTime[c] = now
        for c in self.StopCounters:
            if c in TimingLayer.StartTime:
                TimingLayer.TotalTimeMS[c] = TimingLayer.TotalTimeMS.get(c, 0.0) + (now - TimingLayer.StartTime[c]) * 1000.0
                TimingLayer.N[c] = TimingLayer.N.get(c, 0) + 1
        return m


class SquareActivation(BaseLayer):
    """`NeuralNetworks/SquareActivation.cs:8-20`"""

    def Apply(self, m):
        return m.ElementWiseMultiply(m, self.Factory.AllocateComputationEnv())

    def GetOutputScale(self):
        s = self.Source.GetOutputScale()
        return s * s


class _ConvLayerBase(BaseLayer):
    def __init__(self, **kw):
        self.ce = ConvolutionEngine()
        self.Weights = None
        self.Bias = None
        self.WeightsScale = 1.0
        self.weightWindows = None
        self.biasVectors = None
        self.kernelSize = -1
        super().__init__(**kw)

    InputShape = property(lambda s: s.ce.InputShape, lambda s, v: setattr(s.ce, "InputShape", list(v)))
    KernelShape = property(lambda s: s.ce.KernelShape, lambda s, v: setattr(s.ce, "KernelShape", list(v)))
    Stride = property(lambda s: s.ce.Stride, lambda s, v: setattr(s.ce, "Stride", list(v)))
    Padding = property(lambda s: s.ce.Padding, lambda s, v: setattr(s.ce, "Padding", list(v)))
    Upperpadding = property(lambda s: s.ce.Upperpadding, lambda s, v: setattr(s.ce, "Upperpadding", list(v)))
    Lowerpadding = property(lambda s: s.ce.Lowerpadding, lambda s, v: setattr(s.ce, "Lowerpadding", list(v)))
    MapCount = property(lambda s: s.ce.MapCount, lambda s, v: setattr(s.ce, "MapCount", list(v)))
    Offsets = property(lambda s: s.ce.Offsets)
    Corners = property(lambda s: s.ce.Corners)

    def maps(self):
        return int(np.prod(self.MapCount)) if self.MapCount is not None else 1

    def GetOutputScale(self):
        return (len(self.Offsets) if self.Weights is None else self.WeightsScale) * self.Source.GetOutputScale()

    def _weight(self, offset, bias):
        l = self.ce.Location(None, offset, self.KernelShape, bias)
        return 0.0 if l < 0 else self.Weights[l]

    def PrepareWeightsWindows(self):  # PoolLayer.cs:101-111
        self.weightWindows = []
        for m in range(self.maps()):
            w = [self._weight(off, m * self.kernelSize) for off in self.Offsets]
            self.weightWindows.append(self.Factory.GetPlainVector(np.array(w), EVectorFormat.sparse, self.WeightsScale))

    def _bias_value(self, mapIndex):
        return self.Bias[mapIndex] if self.Bias is not None else self.Weights[(mapIndex + 1) * self.kernelSize - 1]

    def Dispose(self):
        for lst in (self.weightWindows, self.biasVectors):
            if lst:
                for v in lst:
                    if v is not None:
                        v.Dispose()
        self.weightWindows = self.biasVectors = None

    def OutputDimension(self):
        if not self.layerPrepared:
            self.Prepare()
        return len(self.Corners) * (1 if self.Weights is None else self.maps())


class PoolLayer(_ConvLayerBase):
    """Convolution / dense / mean-pool over per-pixel ciphertexts (`NeuralNetworks/PoolLayer.cs:13-245`)."""

    def __init__(self, Fused=True, **kw):
        self.Fused = Fused
        super().__init__(**kw)

    def Prepare(self):
        if self.layerPrepared:
            return
        self.ce.Prepare()
        self.kernelSize = int(np.prod(self.KernelShape)) + (1 if self.Bias is None else 0)
        if self.Weights is None:
            return
        self.PrepareWeightsWindows()
        self.biasVectors = None
        self.layerPrepared = True

    def _gather_row(self, corner):
        return [self.ce.Location(corner, off, self.InputShape) for off in self.Offsets]

    def Apply(self, m):
        f = self.Factory
        env = f.AllocateComputationEnv()
        if self.Weights is None:  # mean pool: sum of the window, scale absorbs 1/len (PoolLayer.cs:124-145)
            outs = []
            for corner in self.Corners:
                agg = None
                for l in self._gather_row(corner):
                    if l < 0:
                        continue
                    el = m.GetColumn(l)
                    nxt = el if agg is None else agg.Add(el, env)
                    if agg is not None and agg is not el and not _is_column(m, agg):
                        agg.Dispose()
                    agg = nxt
                if _is_column(m, agg):
                    agg = f.CopyVector(agg)
                agg.RegisterScale(agg.Scale * len(self.Offsets))
                outs.append(agg)
            return f.GetMatrix(outs, EMatrixFormat.ColumnMajor, CopyVectors=False)
        maps = self.maps()
        if self.biasVectors is None or self.biasVectors[0].Dim != m.RowCount:
            if self.biasVectors:
                for b in self.biasVectors:
                    b.Dispose()
            scale = self.Source.GetOutputScale() * self.WeightsScale
            self.biasVectors = [f.GetPlainVector(np.full(m.RowCount, self._bias_value(k)), EVectorFormat.dense, scale) for k in range(maps)]
        K = len(self.Offsets)
        M = maps * len(self.Corners)
        if self.Fused and hasattr(f, "ConvDenseLayer"):
            if getattr(self, "_gather", None) is None:  # the topology is static: index table built once
                self._gather = np.array([self._gather_row(c) for c in self.Corners] * maps, dtype=np.int32)  # k = map*corners + corner
            gather = self._gather
            inputs = [m.GetColumn(i) for i in range(m.ColumnCount)]
            weights = [self.weightWindows[k // len(self.Corners)] for k in range(M)]
            bias = [self.biasVectors[k // len(self.Corners)] for k in range(M)]
            res = f.ConvDenseLayer(inputs, gather, weights, bias, M, K)
            return f.GetMatrix(res, EMatrixFormat.ColumnMajor, CopyVectors=False)
        res, temps = [], []
        for k in range(M):  # the reference's per-output path (PoolLayer.cs:113-121, 214-223)
            mapIndex, cornerIndex = divmod(k, len(self.Corners))
            cols = []
            for l in self._gather_row(self.Corners[cornerIndex]):
                if l < 0:
                    z = np.zeros(m.RowCount)
                    zv = f.GetEncryptedVector(z, EVectorFormat.dense, m.Scale) if m.IsEncrypted else f.GetPlainVector(z, EVectorFormat.dense, m.Scale)
                    temps.append(zv)
                    cols.append(zv)
                else:
                    cols.append(m.GetColumn(l))
            patch = f.GetMatrix(cols, EMatrixFormat.ColumnMajor, CopyVectors=False)
            patch.DataDisposedExternaly = True
            conv = patch.Mul(self.weightWindows[mapIndex], env)
            res.append(conv.Add(self.biasVectors[mapIndex], env))
            conv.Dispose()
        for t in temps:
            t.Dispose()
        return f.GetMatrix(res, EMatrixFormat.ColumnMajor, CopyVectors=False)


def _is_column(m, v):
    return any(v is c for c in getattr(m, "vectors", []) or [])


class LLPoolLayer(_ConvLayerBase):
    """`NeuralNetworks/LLPoolLayer.cs:10-153`: the input matrix is [corners x offsets] (im2col), one column per offset."""

    def __init__(self, **kw):
        self.HotIndices = None
        super().__init__(**kw)

    def Prepare(self):
        if self.layerPrepared:
            return
        self.ce.Prepare()
        self.kernelSize = int(np.prod(self.KernelShape)) + (1 if self.Bias is None else 0)
        if self.Weights is None:
            return
        self.PrepareWeightsWindows()
        if self.HotIndices is None:
            self.HotIndices = np.ones(len(self.Corners))
        scale = self.Source.GetOutputScale() * self.WeightsScale
        self.biasVectors = [self.Factory.GetPlainVector(self.HotIndices * self._bias_value(k), EVectorFormat.dense, scale) for k in range(self.maps())]
        self.layerPrepared = True

    def Apply(self, m):
        f = self.Factory
        env = f.AllocateComputationEnv()
        if self.Weights is None:
            vec = None
            for i in range(m.ColumnCount):
                c = m.GetColumn(i)
                vec = c if vec is None else vec.Add(c, env)
            vec.RegisterScale(vec.Scale * m.ColumnCount)
            return f.GetMatrix([vec], EMatrixFormat.ColumnMajor, CopyVectors=False)
        res = []
        for k in range(len(self.biasVectors)):
            mul = m.Mul(self.weightWindows[k], env)
            res.append(mul.Add(self.biasVectors[k], env))
            mul.Dispose()
        return f.GetMatrix(res, EMatrixFormat.ColumnMajor, CopyVectors=False)


class LLVectorizeLayer(BaseLayer):
    """`NeuralNetworks/LLVectorizeLayer.cs:8-24`"""

    OutputDim = -1

    def Apply(self, m):
        vec = m.ConvertToColumnVector(self.Factory.AllocateComputationEnv())
        return self.Factory.GetMatrix([vec], EMatrixFormat.ColumnMajor, CopyVectors=False)

    def OutputDimension(self):
        return self.OutputDim if self.OutputDim > 0 else super().OutputDimension()


class LLDenseLayer(BaseLayer):
    """`NeuralNetworks/LLDenseLayer.cs:10-76`: row-major plain weights x one encrypted column vector (rotate-and-sum)."""

    def __init__(self, **kw):
        self.Weights = None
        self.Bias = None
        self.WeightsScale = 1.0
        self.InputFormat = EVectorFormat.dense
        self.ForceDenseFormat = False
        self.WeightsMatrix = None
        self.BiasVector = None
        self.Shard = None  # (rank, world, process group): split the rows of this layer over the ranks of ONE inference
        self._first_row = 0
        super().__init__(**kw)

    def GetOutputScale(self):
        return self.WeightsScale * self.Source.GetOutputScale()

    def Prepare(self):
        if self.layerPrepared:
            return
        if self.ForceDenseFormat and self.InputFormat == EVectorFormat.sparse:
            raise Exception("forcing dense format is only available when the input is dense")
        f = self.Factory
        rows = len(self.Bias)
        w = np.asarray(self.Weights, dtype=np.float64).reshape(rows, -1)
        bscale = self.Source.GetOutputScale() * self.WeightsScale
        if self.Shard is not None and self.InputFormat == EVectorFormat.dense and not isinstance(f, RawFactory):
            from .parallel import row_slice
            self._first_row, count = row_slice(rows, self.Shard[0], self.Shard[1])
            w = w[self._first_row:self._first_row + count]  # this rank encodes and holds its slice of the rows only
        else:
            self.Shard = None
        if self.InputFormat == EVectorFormat.dense:
            self.BiasVector = f.GetPlainVector(np.asarray(self.Bias), EVectorFormat.dense if self.ForceDenseFormat else EVectorFormat.sparse, bscale)
            self.WeightsMatrix = f.GetPlainMatrix(w, EMatrixFormat.RowMajor, self.WeightsScale)
        else:
            self.BiasVector = f.GetPlainVector(np.asarray(self.Bias), EVectorFormat.dense, bscale)
            self.WeightsMatrix = f.GetPlainMatrix(w, EMatrixFormat.ColumnMajor, self.WeightsScale)
        self.layerPrepared = True

    def OutputDimension(self):
        return len(self.Bias)

    def Apply(self, m):
        if m.ColumnCount > 1:
            raise Exception("Expecting only one column")
        env = self.Factory.AllocateComputationEnv()
        if self.Shard is not None:
            from . import parallel
            rank, world, group = self.Shard
            rows = len(self.Bias)
            part = self.WeightsMatrix.MulRows(m.GetColumn(0), self.ForceDenseFormat, self._first_row, rows)
            if self.ForceDenseFormat:  # partial sums at their global columns: all-gather + local modular adds
                mul = parallel.allreduce_ciphertext_sum(self.Factory, part, group)
            else:                      # this rank's sparse elements: concatenate the slices
                mul = parallel.allgather_sparse_elements(self.Factory, part, [parallel.row_slice(rows, r, world)[1] for r in range(world)], group)
            if mul is not part:
                part.Dispose()
        else:
            mul = self.WeightsMatrix.Mul(m.GetColumn(0), env, self.ForceDenseFormat)
        res = mul.Add(self.BiasVector, env)
        mul.Dispose()
        return self.Factory.GetMatrix([res], EMatrixFormat.ColumnMajor, CopyVectors=False)

    def Dispose(self):
        if self.WeightsMatrix is not None:
            self.WeightsMatrix.Dispose()
        if self.BiasVector is not None:
            self.BiasVector.Dispose()
        self.WeightsMatrix = self.BiasVector = None


class LLSingleLineReader(MatrixSource):
    """One image per GetNext() as a single column vector (`LLSingleLineReader.cs`; TSV parsing replaced by an in-memory array)."""

    def __init__(self, features, Scale, NormalizationFactor):
        super().__init__(features, Scale, NormalizationFactor, MaxSlots=1)

    def GetNext(self):
        if self.pos >= len(self.features):
            self.pos = 0
        f = self.features[self.pos] * self.NormalizationFactor
        self.pos += 1
        return RawMatrix(f.reshape(-1, 1), self.Scale, EMatrixFormat.ColumnMajor, 0)


class LLDuplicateLayer(BaseLayer):
    """`NeuralNetworks/LLDuplicateLayer.cs:8-29`: every column is replicated `Count` times at power-of-two strides."""

    Count = 1

    def Apply(self, m):
        env = self.Factory.AllocateComputationEnv()
        cols = [m.GetColumn(i).Duplicate(int(self.Count), env) for i in range(m.ColumnCount)]
        return self.Factory.GetMatrix(cols, m.Format, CopyVectors=False)

    def OutputDimension(self):
        shift, dim = 1, self.Source.OutputDimension()
        while shift < dim:
            shift *= 2
        return shift * int(self.Count)


class LLInterleaveLayer(BaseLayer):
    """`NeuralNetworks/LLInterleaveLayer.cs:12-56`: keep the selected slots of every column (mask), then pack the columns into one
    vector, column c shifted by c*Shift slots."""

    def __init__(self, **kw):
        self.Shift = 0
        self.SelectedIndices = None
        self.InputGrossDimension = -1
        self.mask = None
        super().__init__(**kw)

    def Prepare(self):
        if self.mask is not None:
            return
        if self.InputGrossDimension < 0:
            self.InputGrossDimension = max(self.SelectedIndices) + 1
        hot = np.zeros(self.InputGrossDimension)
        hot[list(self.SelectedIndices)] = 1.0
        self.mask = self.Factory.GetPlainVector(hot, EVectorFormat.dense, 1)

    def Apply(self, m):
        f = self.Factory
        env = f.AllocateComputationEnv()
        clean = [m.GetColumn(i).PointwiseMultiply(self.mask, env) for i in range(m.ColumnCount)]
        cm = f.GetMatrix(clean, EMatrixFormat.ColumnMajor, CopyVectors=False)
        packed = cm.Interleave(self.Shift, env)
        cm.Dispose()
        return f.GetMatrix([packed], EMatrixFormat.ColumnMajor, CopyVectors=False)

    def OutputDimension(self):
        return self.InputGrossDimension

    def Dispose(self):
        if self.mask is not None:
            self.mask.Dispose()
        self.mask = None


class LLPackedDenseLayer(BaseLayer):
    """`NeuralNetworks/LLPackedDenseLayer.cs:10-76`: `PackingCount` weight rows share one plaintext, each in its own
    `PackingShift`-slot segment; one multiply + partial rotate-and-sum evaluates them all against the duplicated input, the
    result of segment c landing in its last slot (where the bias sits)."""

    def __init__(self, **kw):
        self.Weights = None
        self.Bias = None
        self.WeightsScale = 1.0
        self.PackingCount = 1
        self.PackingShift = 0
        self.WeightsMatrix = None
        self.BiasMatrix = None
        super().__init__(**kw)

    def GetOutputScale(self):
        return self.WeightsScale * self.Source.GetOutputScale()

    def Prepare(self):
        if self.layerPrepared:
            return
        maps = len(self.Bias)
        w = np.asarray(self.Weights, dtype=np.float64).reshape(maps, -1)
        pc, ps = int(self.PackingCount), int(self.PackingShift)
        rows = (maps + pc - 1) // pc
        stacked = np.zeros((rows, pc * ps))
        bias = np.zeros((rows, pc * ps))
        for i in range(maps):
            row, col = divmod(i, pc)
            stacked[row, col * ps: col * ps + w.shape[1]] = w[i]
            bias[row, (col + 1) * ps - 1] = self.Bias[i]
        f = self.Factory
        self.BiasMatrix = f.GetPlainMatrix(bias, EMatrixFormat.RowMajor, self.Source.GetOutputScale() * self.WeightsScale)
        self.WeightsMatrix = f.GetPlainMatrix(stacked, EMatrixFormat.RowMajor, self.WeightsScale)
        self.layerPrepared = True

    def OutputDimension(self):
        return len(self.Bias)

    def Apply(self, m):
        if m.ColumnCount > 1:
            raise Exception("Expecting only one column")
        env = self.Factory.AllocateComputationEnv()
        v = m.GetColumn(0)
        res = []
        for k in range(self.WeightsMatrix.RowCount):
            mul = self.WeightsMatrix.GetRow(k).DotProduct(v, env, length=int(self.PackingShift))
            res.append(mul.Add(self.BiasMatrix.GetRow(k), env))
            mul.Dispose()
        return self.Factory.GetMatrix(res, EMatrixFormat.ColumnMajor, CopyVectors=False)

    def Dispose(self):
        for mat in (self.WeightsMatrix, self.BiasMatrix):
            if mat is not None:
                mat.Dispose()
        self.WeightsMatrix = self.BiasMatrix = None


class LLInterleavedDenseLayer(BaseLayer):
    """`NeuralNetworks/LLInterleavedDenseLayer.cs:12-77`: dense layer whose inputs sit at the slots an LLInterleaveLayer left them in."""

    def __init__(self, **kw):
        self.Weights = None
        self.Bias = None
        self.WeightsScale = 1
        self.Shift = 0
        self.SelectedIndices = None
        self.WeightsMatrix = None
        self.BiasVector = None
        super().__init__(**kw)

    def GetOutputScale(self):
        return self.Source.GetOutputScale() * self.WeightsScale

    def OutputDimension(self):
        return len(self.Bias)

    def _targets(self, count):
        out, offset = [], 0
        while count > 0:
            for s in self.SelectedIndices:
                if count == 0:
                    break
                out.append(s + offset)
                count -= 1
            offset += self.Shift
        return out

    def Prepare(self):
        if self.layerPrepared:
            return
        rows = len(self.Bias)
        small = np.asarray(self.Weights, dtype=np.float64).reshape(rows, -1)
        big = np.zeros((rows, self.Source.OutputDimension()))
        for i, t in enumerate(self._targets(small.shape[1])):
            big[:, t] = small[:, i]
        f = self.Factory
        self.BiasVector = f.GetPlainVector(np.asarray(self.Bias, dtype=np.float64), EVectorFormat.sparse, self.GetOutputScale())
        self.WeightsMatrix = f.GetPlainMatrix(big, EMatrixFormat.RowMajor, self.WeightsScale)
        self.layerPrepared = True

    def Apply(self, m):
        env = self.Factory.AllocateComputationEnv()
        mul = self.WeightsMatrix.Mul(m.GetColumn(0), env)
        v = mul.Add(self.BiasVector, env)
        mul.Dispose()
        return self.Factory.GetMatrix([v], EMatrixFormat.ColumnMajor, CopyVectors=False)

    def Dispose(self):
        if self.WeightsMatrix is not None:
            self.WeightsMatrix.Dispose()
        if self.BiasVector is not None:
            self.BiasVector.Dispose()
        self.WeightsMatrix = self.BiasVector = None


class LLPreConvLayer(BaseLayer):
    """`NeuralNetworks/LLPreConvLayer.cs:13-170`: builds the im2col columns of a convolution *homomorphically* from one encrypted
    image vector.  Column i (kernel offset i) is a permutation of the image: for every block of output rows, mask the pixels that
    offset touches and rotate them so that output position `CornersMap[j]` holds corner j's pixel -- the same map for every offset,
    so the following LLPoolLayer is a plain column-wise weighted sum (with `HotIndices` marking the live slots for the bias)."""

    def __init__(self, **kw):
        self.ce = ConvolutionEngine()
        self.UseAxisForBlocks = None
        self.outputDim = -1
        self.shifts = None
        self.masks = None
        self.CornersMap = None
        self._hot = None
        super().__init__(**kw)

    InputShape = property(lambda s: s.ce.InputShape, lambda s, v: setattr(s.ce, "InputShape", list(v)))
    KernelShape = property(lambda s: s.ce.KernelShape, lambda s, v: setattr(s.ce, "KernelShape", list(v)))
    Stride = property(lambda s: s.ce.Stride, lambda s, v: setattr(s.ce, "Stride", list(v)))
    Padding = property(lambda s: s.ce.Padding, lambda s, v: setattr(s.ce, "Padding", list(v)))
    Upperpadding = property(lambda s: s.ce.Upperpadding, lambda s, v: setattr(s.ce, "Upperpadding", list(v)))
    Lowerpadding = property(lambda s: s.ce.Lowerpadding, lambda s, v: setattr(s.ce, "Lowerpadding", list(v)))

    @property
    def HotIndices(self):
        if not self.layerPrepared:
            self.Prepare()
        return self._hot

    def _block_offsets(self):
        """Offsets of the stride cosets used as blocks: an odometer over the flagged axes, axis 0 fastest (:31-59)."""
        n = len(self.Stride)
        step = [1] * n
        for i in range(1, n):
            step[i] = step[i - 1] * self.InputShape[i - 1]
        block, offset, out = [0] * n, 0, []
        while True:
            out.append(offset)
            advanced = False
            for i in range(n):
                if not self.UseAxisForBlocks[i]:
                    continue
                block[i] += 1
                offset += step[i]
                if block[i] < self.Stride[i]:
                    advanced = True
                    break
                offset -= block[i] * step[i]
                block[i] = 0
            if not advanced:
                return out

    def Prepare(self):
        if self.layerPrepared:
            return
        ce = self.ce
        ce.Prepare()
        if self.UseAxisForBlocks is None:
            self.UseAxisForBlocks = [True] * len(self.InputShape)
        dim = int(np.prod(ce.InputShape))
        row = dim // ce.InputShape[0]
        boff = self._block_offsets()
        nb = len(boff)
        first_axis = sorted({c[0] for c in ce.Corners})
        small = len(first_axis) // nb
        large = -(-len(first_axis) // nb)
        n_large = len(first_axis) - nb * small
        cmap = [-1] * len(ce.Corners)
        f = self.Factory
        self.masks, self.shifts = [], []
        for off in ce.Offsets:
            sh = [0] * nb
            for j in range(nb):
                size = small if j > n_large else large  # (sic) `>`: block n_large still counts as large in the shift recurrence (:99)
                sh[j] = ce.Location(None, off, ce.InputShape) if j == 0 else sh[j - 1] + boff[j - 1] - boff[j] + size * ce.Stride[0] * row
            sel = [[] for _ in range(nb)]
            for j, corner in enumerate(ce.Corners):
                loc = ce.Location(corner, off, ce.InputShape)
                cid = (corner[0] - ce.Corners[0][0]) // ce.Stride[0]
                block = cid // large if cid < large * n_large else n_large + (cid - large * n_large) // small
                if loc >= 0:
                    sel[block].append(loc)
                    where = loc - sh[block]
                    if cmap[j] >= 0 and cmap[j] != where:
                        raise Exception("Internal Error")
                    cmap[j] = where
            mk = []
            for s in sel:
                if s:
                    hot = np.zeros(dim)
                    hot[s] = 1.0
                    mk.append(f.GetPlainVector(hot, EVectorFormat.dense, 1))
                else:
                    mk.append(None)
            self.masks.append(mk)
            self.shifts.append(sh)
        large_max = 0 if n_large == 0 else row * (1 + ce.Stride[0] * (large - 1)) + boff[n_large - 1]
        small_max = row * (1 + ce.Stride[0] * (small - 1)) + boff[-1]
        self.outputDim = max(large_max, small_max)
        self.CornersMap = cmap
        self._hot = np.zeros(self.outputDim)
        self._hot[cmap] = 1.0
        self.layerPrepared = True

    def Apply(self, m):
        if m.ColumnCount != 1:
            raise Exception("Expecting only a single column")
        if not self.layerPrepared:
            self.Prepare()
        env = self.Factory.AllocateComputationEnv()
        v = m.GetColumn(0)
        cols = [v.Permute(self.masks[k], self.shifts[k], self.outputDim, env) for k in range(len(self.masks))]
        return self.Factory.GetMatrix(cols, EMatrixFormat.ColumnMajor, CopyVectors=False)

    def OutputDimension(self):
        if not self.layerPrepared:
            self.Prepare()
        return self.outputDim

    def RearrangeWeights(self, weights):
        """Weights of the next dense layer re-indexed from corner order to the slot order this layer produces (:154-168)."""
        if not self.layerPrepared:
            self.Prepare()
        weights = np.asarray(weights, dtype=np.float64)
        nc = len(self.ce.Corners)
        maps = len(weights) // nc
        out = np.zeros(maps * self.outputDim)
        for i in range(maps):
            for j in range(nc):
                out[i * self.outputDim + self.CornersMap[j]] = weights[j + i * nc]
        return out

    def Dispose(self):
        if self.masks:
            for mk in self.masks:
                for v in mk:
                    if v is not None:
                        v.Dispose()
        self.masks = None
