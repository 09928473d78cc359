"""Buffer types: same behaviour as the reference's tests/test_buffer_util.py (argument errors,
index errors and their texts, len/size/tobytes, collection semantics)."""
import struct

import pytest

import python_zstandard_b200 as zstd

ss = struct.Struct("=QQ")


def test_arguments():
    with pytest.raises(TypeError):
        zstd.BufferWithSegments()
    with pytest.raises(TypeError):
        zstd.BufferWithSegments(b"foo")
    with pytest.raises(ValueError, match="segments array size is not a multiple of 16"):
        zstd.BufferWithSegments(b"foo", b"\x00\x00")


def test_invalid_offset():
    with pytest.raises(ValueError, match="offset within segments array references memory"):
        zstd.BufferWithSegments(b"foo", ss.pack(0, 4))


def test_invalid_getitem():
    b = zstd.BufferWithSegments(b"foo", ss.pack(0, 3))
    with pytest.raises(IndexError, match="offset must be non-negative"):
        b[-10]
    with pytest.raises(IndexError, match="offset must be less than 1"):
        b[1]
    with pytest.raises(IndexError, match="offset must be less than 1"):
        b[2]


def test_single_and_multiple():
    b = zstd.BufferWithSegments(b"foo", ss.pack(0, 3))
    assert len(b) == 1 and b.size == 3 and b.tobytes() == b"foo"
    assert len(b[0]) == 3 and b[0].offset == 0 and b[0].tobytes() == b"foo"
    b = zstd.BufferWithSegments(b"foofooxfooxy", b"".join([ss.pack(0, 3), ss.pack(3, 4), ss.pack(7, 5)]))
    assert len(b) == 3 and b.size == 12 and b.tobytes() == b"foofooxfooxy"
    assert [b[i].tobytes() for i in range(3)] == [b"foo", b"foox", b"fooxy"]
    assert bytes(memoryview(b)) == b"foofooxfooxy"
    assert bytes(memoryview(b[1])) == b"foox"
    assert bytes(memoryview(b.segments())) == b"".join([ss.pack(0, 3), ss.pack(3, 4), ss.pack(7, 5)])


def test_collection():
    with pytest.raises(ValueError, match="must pass at least 1 argument"):
        zstd.BufferWithSegmentsCollection()
    with pytest.raises(TypeError, match="arguments must be BufferWithSegments"):
        zstd.BufferWithSegmentsCollection(None)
    with pytest.raises(TypeError, match="arguments must be BufferWithSegments"):
        zstd.BufferWithSegmentsCollection(zstd.BufferWithSegments(b"foo", ss.pack(0, 3)), None)
    with pytest.raises(ValueError, match="ZstdBufferWithSegments cannot be empty"):
        zstd.BufferWithSegmentsCollection(zstd.BufferWithSegments(b"", b""))
    b1 = zstd.BufferWithSegments(b"foo", ss.pack(0, 3))
    b2 = zstd.BufferWithSegments(b"barbaz", b"".join([ss.pack(0, 3), ss.pack(3, 3)]))
    c = zstd.BufferWithSegmentsCollection(b1)
    assert len(c) == 1 and c.size() == 3
    c = zstd.BufferWithSegmentsCollection(b1, b2)
    assert len(c) == 3 and c.size() == 9
    with pytest.raises(IndexError, match="offset must be less than 3"):
        c[3]
    assert [c[i].tobytes() for i in range(3)] == [b"foo", b"bar", b"baz"]


def test_device_buffer_argument_checks_without_a_device():
    """DeviceBufferWithSegments validates like BufferWithSegments before it ever touches the GPU."""
    import struct
    import pytest
    import python_zstandard_b200 as zstd

    class Fake:                      # something that claims to be a 1-D device array of 100 bytes (at a host address)
        def __init__(self, shape=(100,), typestr="|u1", strides=None):
            import ctypes
            self._b = ctypes.create_string_buffer(128)
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ctypes.addressof(self._b), False),
                                             "strides": strides, "version": 3}
    with pytest.raises(TypeError, match="__cuda_array_interface__"):
        zstd.DeviceBufferWithSegments(b"host bytes", struct.pack("=QQ", 0, 3))
    with pytest.raises(TypeError, match="contiguous 1-D"):
        zstd.DeviceBufferWithSegments(Fake(shape=(10, 10)), struct.pack("=QQ", 0, 3))
    with pytest.raises(TypeError, match="contiguous 1-D"):
        zstd.DeviceBufferWithSegments(Fake(strides=(2,)), struct.pack("=QQ", 0, 3))
    with pytest.raises(ValueError, match="segments array size is not a multiple of 16"):
        zstd.DeviceBufferWithSegments(Fake(), b"\x00" * 17)
    with pytest.raises(ValueError, match="references memory outside buffer"):
        zstd.DeviceBufferWithSegments(Fake(), struct.pack("=QQ", 90, 11))
    with pytest.raises(TypeError, match="does not live in device memory"):
        zstd.DeviceBufferWithSegments(Fake(), struct.pack("=QQ", 0, 100))      # a host address (or no device at all)
    assert "device_buffers" in zstd.backend_features
