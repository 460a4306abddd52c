"""The tile hand-out of a sharded frame (csg_shard_tile = the host view of shard_tile_coords / slot_of_macro / slots_per_shard,
csrc/csg_kernel.cuh — the same functions the pruning and frame kernels compile): every traced macro tile goes to exactly one
shard, a shard's tiles get distinct tree slots, and no slot lies outside what csg_upload allocates — for any traced rectangle
(it moves with the camera), frame sizes whose tile count is not a multiple of the GPU count, and both shard modes."""
import random

import pytest

import csg_b200 as g


def hand_out(macro_x, macro_y, rect, mode, count):
    owner = {}
    for rank in range(count):
        _, _, _, n_tiles, n_slots = g.shard_tile(macro_x, macro_y, rect, mode, rank, count, -1)
        slots = set()
        for tile in range(n_tiles):
            mx, my, slot, _, _ = g.shard_tile(macro_x, macro_y, rect, mode, rank, count, tile)
            assert rect[0] <= mx < rect[0] + rect[2] and rect[1] <= my < rect[1] + rect[3]
            assert (mx, my) not in owner, "a macro tile was handed out twice"
            owner[(mx, my)] = rank
            assert 0 <= slot < n_slots, f"slot {slot} of {n_slots}: rank {rank} of {count}, rect {rect}, frame {macro_x}x{macro_y}"
            assert slot not in slots, "two tiles of one shard share a slot"
            slots.add(slot)
            if mode == 1:
                assert my % count == rank      # whole macro-tile rows: contiguous in memory, one PCIe link each
    assert len(owner) == rect[2] * rect[3], "a traced macro tile was not handed out"
    return owner


@pytest.mark.parametrize("mode", [0, 1])
def test_every_tile_once_and_slots_in_range_1080p_on_8_gpus(mode):
    # 1920x1080 = 30 x 34 = 1020 macro tiles: not a multiple of 8 (the case ADVICE r1 found: ranks 4-7 got slot 127 of 127)
    rnd = random.Random(5)
    for _ in range(300):
        w, h = rnd.randint(1, 30), rnd.randint(1, 34)
        x0, y0 = rnd.randint(0, 30 - w), rnd.randint(0, 34 - h)
        hand_out(30, 34, (x0, y0, w, h), mode, 8)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("count", [1, 2, 3, 4, 7, 8])
def test_random_frames_and_rectangles(mode, count):
    rnd = random.Random(100 * count + mode)
    for _ in range(120):
        macro_x, macro_y = rnd.randint(1, 70), rnd.randint(1, 70)
        w, h = rnd.randint(0, macro_x), rnd.randint(0, macro_y)
        x0, y0 = rnd.randint(0, macro_x - w), rnd.randint(0, macro_y - h)
        if w == 0 or h == 0:
            w = h = x0 = y0 = 0
        hand_out(macro_x, macro_y, (x0, y0, w, h), mode, count)


def test_tile_mode_interleaves_finely_and_row_mode_gives_whole_rows():
    own0 = hand_out(60, 68, (10, 20, 33, 25), 0, 8)
    assert [own0[(10 + k, 20)] for k in range(8)] == list(range(8))
    own1 = hand_out(60, 68, (10, 20, 33, 25), 1, 8)
    assert all(own1[(mx, my)] == my % 8 for (mx, my) in own1)


def test_4k_full_frame_matches_the_division_by_multiply_high():
    # 60 x 68 macro tiles: the widths the kernels divide by with a host-computed reciprocal (rm_w < 4096, fewer than 2^20 tiles)
    for count in (1, 2, 4, 8):
        own = hand_out(60, 68, (0, 0, 60, 68), 0, count)
        assert all(own[(mx, my)] == (my * 60 + mx) % count for (mx, my) in own)


def test_a_width_coprime_to_the_gpu_count_spreads_every_shard_over_every_column():
    """Why fill_params widens the traced rectangle of a sharded frame (csg_render.cu): tile j goes to shard j mod N along the
    rows, so with a width that is a multiple of N a shard owns whole columns of the frame (Cheese512 @ 4K: 24 x 46 tiles on
    8 GPUs — the shards' loads then differ by what their columns hold), with a coprime width every row is shifted against the
    one above and every shard has tiles in every column."""
    own24 = hand_out(60, 68, (18, 11, 24, 46), 0, 8)
    for rank in range(8):
        assert {mx for (mx, my), r in own24.items() if r == rank} == {18 + rank, 26 + rank, 34 + rank}
    own25 = hand_out(60, 68, (18, 11, 25, 46), 0, 8)
    for rank in range(8):
        assert {mx for (mx, my), r in own25.items() if r == rank} == set(range(18, 43))
        counts = [sum(1 for (mx, my), r in own25.items() if r == rank and mx == col) for col in range(18, 43)]
        assert max(counts) - min(counts) <= 1      # 46 rows of a column dealt out to 8 shards: 5 or 6 each
