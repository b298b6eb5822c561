"""Host-side frame staging of the long-video driver (endodav_b200/video.py::_Stager): the pinned batch a window batch is
uploaded from must hold exactly the frames ``window_frame_indices`` names (endodav.py:185-199), for sequential schedules
of any batch size, including the clamped tail of the video and ring-buffer reuse."""
import numpy as np
import pytest

from endodav_b200 import video as V


class _Done:
    """stands in for the CUDA event recorded after the H2D copy of a staging buffer"""

    def __init__(self):
        self.waited = 0

    def synchronize(self):
        self.waited += 1


@pytest.mark.parametrize("n_frames,WB,world,rank", [(45, 1, 1, 0), (100, 4, 1, 0), (230, 3, 2, 1), (23, 2, 1, 0), (400, 4, 8, 3)])
def test_stager_batches_hold_the_window_frames(n_frames, WB, world, rank):
    rng = np.random.default_rng(n_frames)
    frames = rng.integers(0, 256, size=(n_frames, 6, 8, 3), dtype=np.uint8)
    nwin = V.num_windows(n_frames)
    mine = V.shard_windows(nwin, rank, world)
    st = V._Stager(frames, n_frames, mine, WB, 6, 8)
    events = []
    for j in range(0, len(mine), WB):
        nb = min(WB, len(mine) - j)
        slot, buf = st.fetch(j, nb)
        want = np.concatenate([frames[V.window_frame_indices(mine[j + b], n_frames)] for b in range(nb)])
        assert buf.shape[0] == nb * V.INFER_LEN
        assert np.array_equal(buf.numpy(), want), (j, nb)
        ev = _Done()
        events.append(ev)
        st.uploaded(slot, ev)
    st.close()
    # a buffer is rewritten only after the upload that read it was waited for: every event of a reused slot was synchronised
    rounds = (len(mine) + WB - 1) // WB
    if rounds > V._Stager.RING:
        assert sum(e.waited > 0 for e in events) >= rounds - V._Stager.RING - 2


def test_stager_out_of_order_request_is_still_correct():
    """The prefetch guesses a sequential schedule; a different request is staged on demand."""
    frames = np.random.default_rng(1).integers(0, 256, size=(120, 4, 4, 3), dtype=np.uint8)
    mine = V.shard_windows(V.num_windows(120), 0, 1)
    st = V._Stager(frames, 120, mine, 2, 4, 4)
    for j, nb in ((0, 2), (3, 1), (1, 2), (2, 2), (4, 1), (0, 1), (4, 2), (2, 2)):
        slot, buf = st.fetch(j, nb)
        want = np.concatenate([frames[V.window_frame_indices(mine[j + b], 120)] for b in range(nb)])
        assert np.array_equal(buf.numpy(), want)
        st.uploaded(slot, _Done())
    st.close()
