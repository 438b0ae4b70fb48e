"""Pin the oracle's Philox4x32-10 and curand-uniform restatement (no GPU).

Known-answer vectors are the Random123 (Salmon et al., SC'11) kat_vectors entries for
philox4x32-10, the algorithm curand's Philox4_32_10 generator implements."""
import numpy as np

from oracle import oracle as O

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        assert O.philox4x32_10(ctr, key) == want


def test_curand_uniform_layout():
    """curand_init(seed, subseq, offset): key=seed, ctr=(offset/4 lo, hi, subseq lo, hi);
    draw i = word (offset%4+i)%4 of block offset/4 + (offset%4+i)/4."""
    seed = 0x1234_5678_9abc_def0
    for subseq, offset in [(0, 0), (5, 4), (123456789012, 40), (7, 6), (2**33 + 1, 2**35 + 8)]:
        for i in range(8):
            blk = offset // 4 + (offset % 4 + i) // 4
            w = O.philox4x32_10((blk & 0xffffffff, blk >> 32, subseq & 0xffffffff, subseq >> 32),
                                (seed & 0xffffffff, seed >> 32))[(offset % 4 + i) % 4]
            want = np.float32(np.float32(w) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33))
            got = np.float32(O.curand_uniform(seed, subseq, offset, i))
            assert got == want, (subseq, offset, i)
            assert 0.0 < got <= 1.0


def test_offset_increment_rounding():
    assert O.philox_offset_increment(256, 2) == 1024
    assert O.philox_offset_increment(602, 8) == 604   # 602 rounded up to a multiple of 4
    assert O.philox_offset_increment(13, 2) == 52
    assert O.philox_offset_increment(1, 8) == 4
