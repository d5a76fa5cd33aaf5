"""common/tf_bundle.py: the tensor-bundle (tf.train.Saver V2) reader / writer, checked against the format's own invariants
(no TensorFlow and no reference-written bundle exists in this container: parity unpinned, see the module docstring)."""
import struct

import numpy as np
import pytest

import tcresnet_b200  # noqa: F401
from tcresnet_b200.common import tf_bundle as B


def test_crc32c_known_answers():
    assert B.crc32c(b"123456789") == 0xE3069283                    # the CRC-32C (Castagnoli) check value
    assert B.crc32c(b"") == 0
    assert B.crc32c(bytes(32)) == 0x8A9136AA                       # RFC 3720 B.4: 32 bytes of zeros
    assert B.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43              # RFC 3720 B.4: 32 bytes of ones
    for v in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert B.unmask_crc(B.mask_crc(v)) == v
    assert B.mask_crc(0) == 0xA282EAD8


def test_varint_and_entry_round_trip():
    for n in (0, 1, 127, 128, 300, 2 ** 31, 2 ** 40 + 7):
        assert B._read_varint(B._varint(n), 0) == (n, len(B._varint(n)))
    e = B._decode_entry(B._encode_entry(1, (9, 1, 48, 48), 4096, 82944, 0xDEADBEEF))
    assert e["dtype"] == 1 and e["shape"] == (9, 1, 48, 48) and e["offset"] == 4096 and e["size"] == 82944 and e["crc32c"] == 0xDEADBEEF
    scalar = B._decode_entry(B._encode_entry(9, (), 0, 8, 1))
    assert scalar["shape"] == () and scalar["offset"] == 0 and scalar["shard_id"] == 0


def test_bundle_round_trip_with_many_blocks(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"TCResNet8/block{i // 7}/conv{i}/weights": rng.standard_normal((9, 1, 4 + i % 5, 8)).astype(np.float32) for i in range(300)}
    tensors["global_step"] = np.asarray(12345, np.int64)
    tensors["TCResNet8/fc/weights/Momentum"] = rng.standard_normal((1, 1, 48, 12)).astype(np.float32)
    prefix = B.write_bundle(tmp_path / "Model-12345", tensors)
    index = (tmp_path / "Model-12345.index").read_bytes()
    assert struct.unpack("<Q", index[-8:])[0] == 0xDB4775248B80FB57 and len(index) > 3 * 4096       # several data blocks
    r = B.BundleReader(prefix)
    assert set(r.get_variable_to_shape_map()) == set(tensors) and r.has_tensor("global_step") and not r.has_tensor("nope")
    assert r.get_variable_to_shape_map()["global_step"] == ()
    for k, v in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    # the data file is the tensors back to back in key order
    data = (tmp_path / "Model-12345.data-00000-of-00001").read_bytes()
    assert len(data) == sum(v.nbytes for v in tensors.values())
    first = sorted(tensors, key=lambda s: s.encode())[0]
    assert data[:tensors[first].nbytes] == tensors[first].tobytes()


def test_corruption_is_detected(tmp_path):
    prefix = B.write_bundle(tmp_path / "m-1", {"a": np.arange(64, dtype=np.float32), "b": np.ones((3, 3), np.float32)})
    data = tmp_path / "m-1.data-00000-of-00001"
    raw = bytearray(data.read_bytes())
    raw[5] ^= 0x40
    data.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        B.BundleReader(prefix).get_tensor("a")
    np.testing.assert_array_equal(B.BundleReader(prefix).get_tensor("b"), np.ones((3, 3), np.float32))
    idx = tmp_path / "m-1.index"
    raw = bytearray(idx.read_bytes())
    raw[3] ^= 0x01
    idx.write_bytes(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        B.BundleReader(prefix)


def test_checkpoint_state_file(tmp_path):
    B.write_checkpoint_state(tmp_path, "Model-20", ["Model-10", "Model-20"])
    assert B.read_checkpoint_state(tmp_path) == "Model-20"
    text = (tmp_path / "checkpoint").read_text()
    assert text.splitlines()[0] == 'model_checkpoint_path: "Model-20"' and 'all_model_checkpoint_paths: "Model-10"' in text
