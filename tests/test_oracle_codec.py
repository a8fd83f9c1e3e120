"""P6 container, CPU side: the oracle's restatement (oracle/kmsc_oracle.c kmsc_o_codec_*)
round-trips and produces the documented bytes (layout: csrc/codec.cu). The reference has no
binary k-mer-set format, so these known-answer bytes are derived from the layout by hand."""
import struct
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "kmer-sets-compression_b200"))
import synth  # noqa: E402


def test_known_answer_bytes(oracle):
    K, N, kb = 3, 2, 2                      # 4 buckets, 4 key bits
    # bucket 0: keys 1, 3; bucket 2: key 0x0F; others empty
    offs = np.array([0, 2, 2, 3, 3], np.int64)
    keys = np.array([1, 3, 15], np.uint16)
    data = oracle.codec_encode(K, N, kb, offs, keys)
    head = struct.pack("<6I3Q", 0x43534D4B, 1, K, N, kb, 1, 3, 4, 3)
    # sizes 2,0,1,0 -> all 1-byte codes (control 0x00) + bytes 02 00 01 00
    # deltas 1,2 | 15 -> control 0x00 + bytes 01 02 0f
    assert data == head + bytes([0x00, 2, 0, 1, 0]) + bytes([0x00, 1, 2, 15])
    assert data[:4] == b"KMSC"


def test_code_lengths(oracle):
    K, N, kb = 23, 14, 4
    offs = np.zeros((1 << N) + 1, np.int64)
    keys = np.array([0xFF, 0xFF + 0x100, 0xFF + 0x100 + 0x10000, 0xFF + 0x100 + 0x10000 + 0x1000000], np.uint32)
    offs[8:] = 4                            # all four keys in bucket 7
    data = oracle.codec_encode(K, N, kb, offs, keys)
    n_sctrl = (1 << N) // 4
    key_part = data[48 + n_sctrl + (1 << N):]
    # deltas 0xFF, 0x100, 0x10000, 0x1000000 -> codes 0,1,2,3 = 0b11100100
    assert key_part[0] == 0b11100100
    assert key_part[1:] == bytes([0xFF, 0x00, 0x01, 0x00, 0x00, 0x01, 0x00, 0x00, 0x00, 0x01])


@pytest.mark.parametrize("K,N,kb", [(15, 14, 2), (23, 14, 4), (31, 14, 8), (19, 10, 4), (5, 3, 2)])
def test_roundtrip(oracle, K, N, kb):
    km = synth.kmer_set_of(synth.random_genome(5000, K), K)
    offs, keys = synth.csr_of(km, K, N, kb)
    data = oracle.codec_encode(K, N, kb, offs, keys)
    K2, N2, kb2, offs2, keys2 = oracle.codec_decode(data)
    assert (K2, N2, kb2) == (K, N, kb)
    assert np.array_equal(offs2, offs) and np.array_equal(keys2, keys)


def test_corruption_detected(oracle):
    km = synth.kmer_set_of(synth.random_genome(300, 15), 15)
    offs, keys = synth.csr_of(km, 15, 14, 2)
    data = bytearray(oracle.codec_encode(15, 14, 2, offs, keys))
    with pytest.raises(ValueError):
        oracle.codec_decode(bytes(data[:-1]))
    bad = bytearray(data)
    bad[0] ^= 1
    with pytest.raises(ValueError):
        oracle.codec_decode(bytes(bad))
