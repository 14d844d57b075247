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


def test_pauli_units_round_trip_matches_the_oracle():
    """The boundary form of a Pauli string (units mask: bit 3 s + t <-> site s carries type t + 1) is the same in the product's host
    helper and in the oracle port, and converts back exactly (PauliString::network_unit_at, include/basis/PauliString.hpp:84-90)."""
    import numpy as np
    from annongpu_b200.api import paulis_to_units, units_to_paulis
    from oracle import port_oracle as P
    rng = np.random.default_rng(1)
    for ns in (1, 3, 21, 22, 40, 64, 85):
        for _ in range(8):
            a, b = int(rng.integers(0, 1 << min(ns, 62))), int(rng.integers(0, 1 << min(ns, 62)))
            u = paulis_to_units(a, b, ns)
            assert u.shape == ((3 * ns + 63) // 64,)
            assert units_to_paulis(u, ns) == (a, b)
            words = (ns + 63) // 64
            pa = np.array([(a >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(words)], dtype=np.uint64)
            pb = np.array([(b >> (64 * w)) & 0xFFFFFFFFFFFFFFFF for w in range(words)], dtype=np.uint64)
            assert np.array_equal(u, P.paulis_to_units(pa, pb, ns))
            # at most one unit per site
            v = sum(int(x) << (64 * w) for w, x in enumerate(u))
            assert all(bin((v >> (3 * s)) & 7).count("1") <= 1 for s in range(ns))
