"""Host-side helpers of the binding mirror that need no GPU: Spins (include/basis/Spins.h semantics), mask packing, sharding."""
import numpy as np

from annongpu_b200 import api
from annongpu_b200.distributed import shard_range
from annongpu_b200.factories import masks_to_words, words_for


def test_spins_semantics():
    s = api.Spins(0b1011, 6)
    assert list(s.array()) == [1.0, 1.0, -1.0, 1.0, -1.0, -1.0]          # bit i <-> site i, 1 <-> +1 (Spins.h:104-108)
    assert s.flip(2).configuration == 0b1111 and s.flip(2).flip(2) == s
    assert s.roll(2, 6).configuration == 0b101100                          # Spins::roll (Spins.h:349-354): cyclic left shift
    assert s.roll(6, 6) == s
    big = api.Spins((1 << 130) | 5, 200)
    assert list(big.words()) == [5, 0, 4, 0] and big.words().dtype == np.uint64


def test_mask_packing_and_word_count():
    assert [words_for(n) for n in (1, 64, 65, 128, 129, 256)] == [1, 1, 2, 2, 3, 4]
    w = masks_to_words([(1 << 64) | 3, 1 << 127], 2)
    assert w.tolist() == [[3, 1], [0, 1 << 63]]


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 8192, 131072):
        for world in (1, 2, 3, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert sum(n for _, n in parts) == total
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1)) and parts[0][0] == 0
