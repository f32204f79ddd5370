"""GPU: K0, the on-device FITS pixel decode (csrc/tfm_fits.cu), against the host reader's NumPy path
(fitswcs.read_fits), bit for bit, for every BITPIX, with and without BSCALE/BZERO, ragged sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

_DT = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def write_fits(path, data, bitpix, extra=()):
    """Minimal single-HDU FITS writer (80-character cards, 2880-byte blocks, big-endian data)."""
    cards = [("SIMPLE", "T"), ("BITPIX", str(bitpix)), ("NAXIS", str(data.ndim))]
    for i, s in enumerate(data.shape[::-1]):
        cards.append(("NAXIS%d" % (i + 1), str(s)))
    cards += list(extra)
    txt = "".join(("%-8s= %20s" % (k, v)).ljust(80) for k, v in cards) + "END".ljust(80)
    txt = txt.ljust(-(-len(txt) // 2880) * 2880)
    raw = np.ascontiguousarray(data, dtype=_DT[bitpix]).tobytes()
    raw += b"\0" * (-len(raw) % 2880)
    with open(path, "wb") as f:
        f.write(txt.encode("ascii") + raw)


def _samples(rng, bitpix, shape):
    if bitpix == 8:
        return rng.integers(0, 256, shape)
    if bitpix == 16:
        return rng.integers(-32768, 32768, shape)
    if bitpix == 32:
        return rng.integers(-2 ** 31, 2 ** 31, shape)
    if bitpix == 64:
        return rng.integers(-2 ** 52, 2 ** 52, shape)
    a = (rng.random(shape) - 0.5) * 1e4
    a.flat[0] = np.nan
    a.flat[1] = -0.0
    return a


@pytest.mark.parametrize("bitpix", [8, 16, 32, 64, -32, -64])
@pytest.mark.parametrize("shape", [(37, 53), (3, 17, 29), (1, 5)])
def test_decode_matches_host_reader(tfm, tmp_path, bitpix, shape):
    rng = np.random.default_rng(abs(bitpix) + len(shape))
    data = _samples(rng, bitpix, shape)
    for extra in ((), (("BSCALE", "0.25"), ("BZERO", "32768.5"))):
        path = str(tmp_path / ("f%d_%d.fits" % (bitpix, len(extra))))
        write_fits(path, data, bitpix, extra)
        host = tfm.read_fits(path)[0].data
        devd = tfm.read_fits(path, device=True)[0].data
        assert devd.is_cuda and tuple(devd.shape) == host.shape
        got = devd.cpu().numpy()
        if bitpix == -32:
            assert got.dtype == np.float32 == host.dtype
            if not extra:
                assert np.array_equal(got.view(np.uint32), host.view(np.uint32))  # a pure byte swap
            else:
                assert np.array_equal(got, host, equal_nan=True)                  # float32 scaling
        else:
            assert got.dtype == np.float64
            assert np.array_equal(got, host.astype(np.float64), equal_nan=True)


def test_decoded_frame_feeds_resample(tfm, oracle, tmp_path):
    from conftest import to_oracle
    rng = np.random.default_rng(5)
    data = rng.random((64, 80)).astype(np.float32)
    path = str(tmp_path / "img.fits")
    write_fits(path, data, -32)
    frame = tfm.read_fits(path, device=True)[0].data
    tr = tfm.Composition([tfm.Scale(1.2, dim=2), tfm.Rotation(10, unit='deg'), tfm.Offset([-40.0, -32.0])])
    got = tr.resample(frame, method='linear')
    want = oracle.resample(to_oracle(tr), data, method='l')
    assert np.array_equal(got.cpu().numpy(), want)


def test_decode_throughput_smoke(tfm):
    """Large unscaled float32 unit: round trip through the byte swap twice is the identity."""
    import torch
    n = 1 << 24
    a = torch.rand(n, device="cuda", dtype=torch.float32)
    raw = a.view(torch.uint8)
    once = tfm.decode_fits_data(raw, -32, (n,))
    twice = tfm.decode_fits_data(once.view(torch.uint8), -32, (n,))
    assert torch.equal(twice, a)
    with pytest.raises(ValueError):
        tfm.decode_fits_data(raw[: n], -32, (n,))             # too few bytes
