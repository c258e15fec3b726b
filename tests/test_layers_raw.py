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
