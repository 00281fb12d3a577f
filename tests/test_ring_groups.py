"""The streaming attention (csrc/attn_mma.cu) fetches only the 8-key groups of a stream's 288-slot K / V ring that hold a valid key.
Which groups those are is pure integer algebra on (ring head, cache length, new rows): the product computes it in closed form
(csrc/common.cuh ring_valid_groups, two shifts and masks), exported through the C ABI as pkb_debug_ring_valid_groups so that it can be
checked here, without a GPU, against the definition -- slot p is valid iff its logical position j = (p - head) mod 288 lies in
[256 - len, 256 + qlen)  (NeMo's cache is a FIFO whose valid region is the SUFFIX of length len, followed by the chunk's own rows;
SURVEY.md section 8a).  Exhaustive over every head and cache length for the chunk sizes the engine produces."""
import numpy as np

import binding

RING, CACHE = 288, 256


def brute(head, ln, qlen):
    p = np.arange(RING)
    j = (p - head) % RING
    valid = (j >= CACHE - ln) & (j < CACHE + qlen)
    return valid.reshape(RING // 8, 8).any(axis=1)


def test_group_mask_equals_definition_exhaustively():
    lib = binding.load_library()
    checked = 0
    for qlen in (1, 3, 4, 6, 8, 16, 32):
        for head in range(RING):
            for ln in (list(range(0, 40)) + list(range(40, 257, 7)) + [254, 255, 256]):
                got = lib.pkb_debug_ring_valid_groups(head, ln, qlen)
                want = brute(head, ln, qlen)
                bits = np.array([(got >> q) & 1 for q in range(RING // 8)], bool)
                assert (bits == want).all(), (head, ln, qlen, bin(got))
                checked += 1
    assert checked > 100000


def test_steady_state_fetch_is_within_3_percent_of_the_valid_rows():
    """Saturated cache, 6 new rows: 262 valid slots; the head advances by 3 per chunk, so over a full ring period the fetched groups
    average 269 slots = 2.7 % above the valid rows (whole 96-key blocks: 288 slots = 9.9 % above)."""
    lib = binding.load_library()
    slots = [8 * bin(lib.pkb_debug_ring_valid_groups(h, 256, 6)).count("1") for h in range(0, RING, 3)]
    assert min(slots) >= 262 and max(slots) <= 272
    assert np.mean(slots) / 262.0 < 1.03


def test_out_of_geometry_arguments_return_zero():
    lib = binding.load_library()
    assert lib.pkb_debug_ring_valid_groups(288, 10, 6) == 0 and lib.pkb_debug_ring_valid_groups(-1, 10, 6) == 0
    assert lib.pkb_debug_ring_valid_groups(0, 257, 6) == 0 and lib.pkb_debug_ring_valid_groups(0, 0, 33) == 0
    # an empty cache and no new rows: the kernel never runs such an entry; the function still names the group of the run's start
    assert bin(lib.pkb_debug_ring_valid_groups(5, 0, 0)).count("1") == 1
