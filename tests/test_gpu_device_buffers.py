"""Device-resident I/O of the batch calls: DeviceBufferWithSegments in, DeviceBufferWithSegments
out; the bytes never cross PCIe, the results equal the host path's and the reference's."""
import numpy as np
import pytest
import torch

import corpus
import python_zstandard_b200 as zstd
from oracle import RefZstd, have_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason="oracle/_ref is built from /root/reference")]


def _table(lens):
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    return np.stack([off, np.asarray(lens, dtype=np.uint64)], axis=1).astype(np.uint64)


def test_device_resident_round_trip_equals_the_host_path():
    ref = RefZstd()
    blob, off, ln = corpus.silesia_mix(300, 131072)
    segs = np.stack([off, ln], axis=1).astype(np.uint64)
    d_in = torch.from_numpy(blob.copy()).cuda()
    dbuf = zstd.DeviceBufferWithSegments(d_in, segs.tobytes())
    assert len(dbuf) == 300 and dbuf.size == len(blob) and dbuf.device == 0
    c = zstd.ZstdCompressor(level=3)
    dcomp = c.multi_compress_to_buffer(dbuf)                       # frames stay on the device
    assert isinstance(dcomp, zstd.DeviceBufferWithSegments) and len(dcomp) == 300
    host = c.multi_compress_to_buffer(zstd.BufferWithSegments(blob.tobytes(), segs.tobytes()))
    assert [dcomp[i].tobytes() for i in (0, 7, 299)] == [host[i].tobytes() for i in (0, 7, 299)]
    assert dcomp.tobytes() == b"".join(host[i].tobytes() for i in range(300))
    # the reference decodes what the device holds
    assert ref.decompress(dcomp[5].tobytes(), 131072) == blob[off[5]:off[5] + ln[5]].tobytes()
    # device -> device decode, viewed as a torch tensor without a copy
    dout = zstd.ZstdDecompressor().multi_decompress_to_buffer(dcomp)
    assert isinstance(dout, zstd.DeviceBufferWithSegments) and dout.size == len(blob)
    t = torch.as_tensor(dout, device="cuda")
    assert t.data_ptr() == dout.__cuda_array_interface__["data"][0]
    assert torch.equal(t, d_in)
    assert dout[17].tobytes() == blob[off[17]:off[17] + ln[17]].tobytes() and dout[17].offset == off[17]
    hb = dout.to_host()
    assert isinstance(hb, zstd.BufferWithSegments) and hb[299].tobytes() == blob[off[299]:].tobytes()


def test_device_input_made_by_the_reference_with_sizes_and_errors():
    ref = RefZstd()
    text = corpus.text_corpus(1 << 20)
    items = [text[i * 3000:i * 3000 + 2000 + 7 * i].tobytes() for i in range(64)]
    frames = [ref.compress(s, level=3, content_size=bool(i % 2)) for i, s in enumerate(items)]
    d = zstd.ZstdDecompressor()
    tab = _table([len(f) for f in frames])
    dev = zstd.DeviceBufferWithSegments(torch.frombuffer(bytearray(b"".join(frames)), dtype=torch.uint8).cuda(), tab.tobytes())
    sizes = np.array([len(s) for s in items], dtype=np.uint64)
    out = d.multi_decompress_to_buffer(dev, decompressed_sizes=sizes.tobytes())
    assert [out[i].tobytes() for i in range(64)] == items
    with pytest.raises(ValueError, match="could not determine decompressed size of item 0"):
        d.multi_decompress_to_buffer(dev)                                    # item 0 has no content size in its header
    bad = bytearray(b"".join(frames)); bad[tab[10, 0] + 9] ^= 0xFF
    devbad = zstd.DeviceBufferWithSegments(torch.frombuffer(bad, dtype=torch.uint8).cuda(), tab.tobytes())
    with pytest.raises(zstd.ZstdError, match="error decompressing item 10"):
        d.multi_decompress_to_buffer(devbad, decompressed_sizes=sizes.tobytes())
    with pytest.raises(TypeError):
        zstd.DeviceBufferWithSegments(b"host bytes", tab.tobytes())
    with pytest.raises(ValueError, match="references memory outside buffer"):
        zstd.DeviceBufferWithSegments(torch.zeros(10, dtype=torch.uint8, device="cuda"), _table([11]).tobytes())
