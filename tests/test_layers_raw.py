"""Layer library on the Raw backend, restating `NeuralNetworksTest/LayersTest.cs` (EvenPool :53-82) and checking the
CryptoNets-MNIST / LoLa-small topologies against a direct numpy evaluation.  CPU only."""
import numpy as np

from cryptonets_b200.interfaces import EMatrixFormat
from cryptonets_b200.layers import ConvolutionEngine, MatrixSource, PoolLayer
from cryptonets_b200.networks import cryptonets_mnist, cryptonets_weights, lola_small, lola_small_weights, synthetic_mnist, transpose
from cryptonets_b200.raw import RawFactory


def test_even_pool_known_answer():
    # LayersTest.cs:53-82: 3 maps of 4x4, mean pool 2x2 stride 2 -> expected means
    vals = np.arange(1, 49, dtype=np.float64).reshape(1, 48)
    src = MatrixSource(vals, Scale=1.0)
    pool = PoolLayer(Source=src, InputShape=[3, 4, 4], KernelShape=[1, 2, 2], Stride=[1, 2, 2], Factory=RawFactory(8192))
    pool.PrepareNetwork()
    out = pool.GetNext().Decrypt(None)
    expect = [3.5, 5.5, 11.5, 13.5, 19.5, 21.5, 27.5, 29.5, 35.5, 37.5, 43.5, 45.5]
    assert np.allclose(out.reshape(-1), expect)


def test_convolution_engine_shapes():
    ce = ConvolutionEngine()
    ce.InputShape, ce.KernelShape, ce.Stride, ce.Upperpadding, ce.MapCount = [28, 28], [5, 5], [2, 2], [1, 1], [5, 1]
    ce.Prepare()
    assert len(ce.Corners) == 169 and len(ce.Offsets) == 25
    pads = sum(1 for c in ce.Corners for o in ce.Offsets if ce.Location(c, o, ce.InputShape) < 0)
    assert pads == 129  # 129 out-of-range (corner, offset) pairs
    assert ce.Offsets[1] == [1, 0] and ce.Corners[1] == [0, 2]  # first axis fastest / last axis fastest


def _numpy_cryptonets(images, w):
    x = np.rint(images / 256.0 * 16.0)
    ce = ConvolutionEngine()
    ce.InputShape, ce.KernelShape, ce.Stride, ce.Upperpadding, ce.MapCount = [28, 28], [5, 5], [2, 2], [1, 1], [5, 1]
    ce.Prepare()
    w0 = np.rint(w["Weights_0"] * 32)
    conv = np.zeros((len(x), 845))
    for m in range(5):
        for ci, c in enumerate(ce.Corners):
            # bias = last weight of the window, at the layer's output scale
            acc = np.rint(np.full(len(x), w["Weights_0"][(m + 1) * 26 - 1] * 16.0 * 32))
            for oi, o in enumerate(ce.Offsets):
                l = ce.Location(c, o, ce.InputShape)
                if l >= 0:
                    acc = acc + w0[m * 26 + ce.Location(None, o, ce.KernelShape)] * x[:, l]
            conv[:, m * 169 + ci] = acc
    a = conv * conv
    s = (16.0 * 32) ** 2
    w1 = np.rint(transpose(w["Weights_1"], 845, 100) * 1024).reshape(100, 845)
    d = a @ w1.T + np.rint(w["Biases_2"] * s * 1024)
    a2 = d * d
    s2 = (s * 1024) ** 2
    w3 = np.rint(w["Weights_3"] * 32).reshape(10, 100)
    out = a2 @ w3.T + np.rint(w["Biases_3"] * s2 * 32)
    return out / (s2 * 32)


def test_cryptonets_topology_on_raw_backend():
    imgs = synthetic_mnist(6, seed=1)
    w = cryptonets_weights()
    net, _ = cryptonets_mnist(RawFactory(8192), imgs, timing=False, weights=w)
    net.PrepareNetwork()
    got = net.GetNext().Decrypt(None)
    assert got.shape == (6, 10)
    want = _numpy_cryptonets(imgs, w)
    assert np.allclose(got, want, rtol=1e-12, atol=0)


def test_lola_small_topology_on_raw_backend():
    imgs = synthetic_mnist(2, seed=2)
    w = lola_small_weights()
    net, _ = lola_small(RawFactory(8192), imgs, weights=w)
    net.PrepareNetwork()
    got = net.GetNext().Decrypt(None).reshape(-1)
    assert got.shape == (10,)
    # direct evaluation: conv (5 maps x 169 corners) -> square -> dense 845 -> 10
    x = np.rint(imgs[0] / 256.0 * 16.0)
    ce = ConvolutionEngine()
    ce.InputShape, ce.KernelShape, ce.Stride, ce.Upperpadding = [28, 28], [5, 5], [2, 2], [1, 1]
    ce.Prepare()
    w0 = np.rint(w["Weights_0"] * 64)
    conv = np.zeros(845)
    for m in range(5):
        for ci, c in enumerate(ce.Corners):
            acc = np.rint(w["Weights_0"][(m + 1) * 26 - 1] * 16.0 * 64)
            for o in ce.Offsets:
                l = ce.Location(c, o, ce.InputShape)
                if l >= 0:
                    acc += w0[m * 26 + ce.Location(None, o, ce.KernelShape)] * x[l]
            conv[m * 169 + ci] = acc
    a = conv * conv
    s = (16.0 * 64) ** 2
    w1 = np.rint(w["Weights_1"] * 64).reshape(10, 845)
    want = (w1 @ a + np.rint(w["Biases_1"] * s * 64)) / (s * 64)
    assert np.allclose(got, want, rtol=1e-12, atol=0)


def test_lola_and_lola_dense_topologies_equal_cryptonets_on_raw_backend():
    """LoLa (LoLaCryptonets.cs:203-276) and LoLa-Dense (:116-201) evaluate the SAME model as CryptoNets (same shipped weights) through
    completely different data flows (im2col + packed dense + interleave; LoLa-Dense additionally builds the im2col columns with
    LLPreConvLayer permutations): all three must give the same scores -- which exercises LLDuplicateLayer, LLPackedDenseLayer,
    LLInterleaveLayer, LLInterleavedDenseLayer, LLPreConvLayer.RearrangeWeights/HotIndices end to end."""
    from cryptonets_b200.networks import lola, lola_dense
    imgs = synthetic_mnist(3, seed=11)
    w = cryptonets_weights()
    want = _numpy_cryptonets(imgs, w)
    for build, block in ((lola, 8192), (lola_dense, 16384)):
        net, _ = build(RawFactory(block), imgs, weights=w)
        net.PrepareNetwork()
        for i in range(len(imgs)):
            got = np.asarray(net.GetNext().Decrypt(None)).reshape(-1)
            assert got.shape == (10,)
            assert np.allclose(got, want[i], rtol=1e-9, atol=1e-9), (build.__name__, i)


def test_preconv_layer_builds_im2col_columns():
    """LLPreConvLayer output column k, slot CornersMap[j] == pixel (corner j + offset k) of the image, 0 where padded (LLPreConvLayer.cs:75-127)."""
    from cryptonets_b200.layers import LLPreConvLayer, LLSingleLineReader
    img = synthetic_mnist(1, seed=3)
    rd = LLSingleLineReader(img, Scale=1.0, NormalizationFactor=1.0)
    from cryptonets_b200.layers import EncryptLayer
    enc = EncryptLayer(Source=rd, Factory=RawFactory(16384))
    pre = LLPreConvLayer(Source=enc, InputShape=[28, 28], KernelShape=[5, 5], Upperpadding=[1, 1], Stride=[2, 2])
    pre.PrepareNetwork()
    out = pre.Apply(enc.GetNext())
    assert out.ColumnCount == 25
    ce = pre.ce
    assert pre.HotIndices.sum() == len(ce.Corners) == 169 and pre.OutputDimension() == len(pre.HotIndices)
    cols = [np.asarray(out.GetColumn(k).Decrypt(None)).reshape(-1) for k in range(25)]
    for k, off in enumerate(ce.Offsets):
        for j, corner in enumerate(ce.Corners):
            l = ce.Location(corner, off, ce.InputShape)
            assert cols[k][pre.CornersMap[j]] == (img[0][l] if l >= 0 else 0)


def test_packed_dense_and_interleave_known_answer():
    """4 outputs packed 2 per plaintext at stride 8 over a duplicated input; interleave gathers the 2 segment results of each row."""
    from cryptonets_b200.interfaces import EMatrixFormat, EVectorFormat
    from cryptonets_b200.layers import LLDuplicateLayer, LLInterleaveLayer, LLPackedDenseLayer
    f = RawFactory(64)
    x = np.arange(1, 7, dtype=np.float64)  # 6 inputs -> padded to 8 by Duplicate

    class Src:
        Factory = f

        def GetOutputScale(self):
            return 1.0

        def OutputDimension(self):
            return 6

        def PrepareNetwork(self):
            pass

    m = f.GetMatrix([f.GetPlainVector(x, EVectorFormat.dense, 1.0)], EMatrixFormat.ColumnMajor, CopyVectors=False)
    dup = LLDuplicateLayer(Source=Src(), Count=2)
    assert dup.OutputDimension() == 16
    d = dup.Apply(m)
    W = np.arange(24, dtype=np.float64).reshape(4, 6) - 7
    b = np.array([0.5, -1.0, 2.0, 3.0])
    dense = LLPackedDenseLayer(Source=dup, Weights=W.reshape(-1), Bias=b, WeightsScale=2.0, PackingCount=2, PackingShift=8)
    dense.Prepare()
    y = dense.Apply(d)
    assert y.ColumnCount == 2
    want = W @ x + b
    got = [np.asarray(y.GetColumn(r).Decrypt(None)).reshape(-1) for r in range(2)]
    for i in range(4):
        row, col = divmod(i, 2)
        assert np.isclose(got[row][(col + 1) * 8 - 1], want[i])
    inter = LLInterleaveLayer(Source=dense, Shift=-1, SelectedIndices=[7, 15])
    inter.Prepare()
    z = np.asarray(inter.Apply(y).GetColumn(0).Decrypt(None)).reshape(-1)
    # column c is shifted by c * Shift = -c slots: output i = 2*row + col sits at slot (col+1)*8 - 1 - row
    for i in range(4):
        row, col = divmod(i, 2)
        assert np.isclose(z[(col + 1) * 8 - 1 - row], want[i])


def test_lola_large_and_cifar_topologies_on_raw_backend():
    """LoLa-Large and LoLa-CIFAR (synthetic weights of the shipped shapes) against a direct numpy evaluation: conv -> square ->
    conv-as-dense (ConvolutionEngine.GetDenseWeights) -> square -> dense."""
    from cryptonets_b200.networks import cifar_weights, lola_cifar, lola_large, lola_large_weights, synthetic_cifar

    def direct(x, w, shape1, kern1, pad, maps1, s0, ws0, shape2, kern2, pad2, stride2, maps2, ws1, ws2):
        ce = ConvolutionEngine()
        ce.InputShape, ce.KernelShape, ce.Stride, ce.MapCount = shape1, kern1, [1000, 2, 2], [maps1, 1, 1]
        ce.Upperpadding, ce.Lowerpadding = pad, pad
        a = ce.GetDenseWeights(np.rint(np.asarray(w["Weights_0"]) * ws0)).reshape(-1, int(np.prod(shape1))) @ x
        a = a + np.rint(ce.GetDenseBias(w["Biases_0"]) * s0 * ws0)
        a = a * a
        ce2 = ConvolutionEngine()
        ce2.InputShape, ce2.KernelShape, ce2.Stride, ce2.MapCount = shape2, kern2, stride2, [maps2, 1, 1]
        if pad2:
            ce2.Upperpadding, ce2.Lowerpadding = pad2, pad2
        s1 = (s0 * ws0) ** 2
        b = np.rint(ce2.GetDenseWeights(w["Weights_1"]) * ws1).reshape(-1, a.size) @ a + np.rint(ce2.GetDenseBias(w["Biases_1"]) * s1 * ws1)
        b = b * b
        s2 = (s1 * ws1) ** 2
        out = np.rint(np.asarray(w["Weights_2"]) * ws2).reshape(10, -1) @ b + np.rint(np.asarray(w["Biases_2"]) * s2 * ws2)
        return out / (s2 * ws2)

    img = synthetic_mnist(1, seed=3)
    w = lola_large_weights()
    net, _ = lola_large(RawFactory(16384), img, weights=w)
    net.PrepareNetwork()
    got = np.asarray(net.GetNext().Decrypt(None)).reshape(-1)
    wl = dict(w, Weights_0=np.asarray(w["Weights_0"]) / 256.0)
    want = direct(np.rint(img[0] * 16.0), wl, [1, 28, 28], [1, 8, 8], [0, 1, 1], 83, 16.0, 4096, [83, 12, 12], [83, 6, 6], None, [83, 2, 2], 163, 64, 512)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
    img = synthetic_cifar(1)
    w = cifar_weights()
    net, _ = lola_cifar(RawFactory(16384), img, weights=w)
    net.PrepareNetwork()
    got = np.asarray(net.GetNext().Decrypt(None)).reshape(-1)
    want = direct(np.rint(img[0] / 256.0 * 8.0), w, [3, 32, 32], [3, 8, 8], [0, 1, 1], 83, 8.0, 256.0, [83, 14, 14], [83, 10, 10], [0, 4, 4], [83, 2, 2], 112,
                  512.0, 512.0)
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12)
