"""Host logic of `video_dqn_b200.valuemap.build_value_maps` (visualize_value.py:60-99 around the forward-only
runner) with a stand-in runner on synthetic view folders: cell discovery, per-orientation view selection /
panorama rotation, padding of the last batch, the scatter into the maps.  CPU only."""
import os

import numpy as np
import torch
from PIL import Image

from video_dqn_b200.valuemap import build_value_maps, list_cells


class FakeRunner:
    """value[b, c] = mean of frame (or of heading c % F of a panorama) -- distinguishes views and their order"""

    def __init__(self, B, panorama):
        self.B, self.panorama, self.calls = B, panorama, 0

    def __call__(self, frames):
        self.calls += 1
        assert frames.dtype == torch.uint8 and frames.shape[0] == self.B and frames.is_contiguous()
        if self.panorama:
            assert frames.shape[1:] == (4, 224, 224, 3)
            m = frames.float().mean(dim=(2, 3, 4))                      # [B, 4]: one number per heading slot
            value = torch.stack([m[:, c % 4] * (c + 1) for c in range(5)], dim=1)
        else:
            assert frames.shape[1:] == (224, 224, 3)
            m = frames.float().mean(dim=(1, 2, 3))
            value = torch.stack([m * (c + 1) for c in range(5)], dim=1)
        return None, value, None


def _render(folder, cells):
    os.makedirs(folder, exist_ok=True)
    level = {}
    for k, (r, c) in enumerate(cells):
        for i in range(4):
            v = 10 + 40 * i + 3 * k                                       # flat grey: survives JPEG + resize exactly
            Image.fromarray(np.full((240, 320, 3), v, np.uint8)).save(os.path.join(folder, f"{r}-{c}-{i}.jpg"), quality=95)
            level[(r, c, i)] = v
    open(os.path.join(folder, "info.npy"), "wb").close()                # non-view files are ignored
    return level


def test_single_view_maps(tmp_path):
    cells = [(3, 7), (3, 8), (10, 2), (11, 2), (0, 0)]
    level = _render(str(tmp_path), cells)
    assert list_cells(str(tmp_path)) == sorted(cells)
    r = FakeRunner(2, panorama=False)                                     # 5 cells: two full batches + a padded one
    maps, free = build_value_maps(str(tmp_path), r, panorama=False, resolution=16, workers=2)
    assert r.calls == 3 * 4 and len(maps) == 4 and maps[0].shape == (16, 16, 5)
    assert free.sum() == len(cells) and all(free[rr, cc] == 1 for rr, cc in cells)
    for ori in range(4):
        for rr, cc in cells:
            np.testing.assert_allclose(maps[ori][rr, cc], [level[(rr, cc, ori)] * (k + 1) for k in range(5)], atol=1e-4)
        assert np.count_nonzero(maps[ori].sum(-1)) == len(cells)          # nothing else written


def test_panorama_rotation(tmp_path):
    cells = [(1, 1), (2, 5), (4, 4)]
    level = _render(str(tmp_path), cells)
    r = FakeRunner(4, panorama=True)
    maps, _ = build_value_maps(str(tmp_path), r, panorama=True, resolution=8, workers=2)
    assert r.calls == 4
    for ori in range(4):
        for rr, cc in cells:
            # heading slot s of the rotated panorama holds view (ori + s) % 4 (cat(images[ori:], images[:ori]))
            want = [level[(rr, cc, (ori + (k % 4)) % 4)] * (k + 1) for k in range(5)]
            np.testing.assert_allclose(maps[ori][rr, cc], want, atol=1e-4)


import pytest  # noqa: E402

REF = os.environ.get("VDQN_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dataloaders")), reason="needs the reference checkout")
def test_views_match_the_reference_dataset_class(tmp_path):
    """The frames `build_value_maps` hands to the runner, normalised as the first kernel does
    (`to_imgnet`), equal what the reference's own `HabitatQVisualizationDatasetGibson` yields for the same
    folder, cell, orientation and panorama switch (dataloaders/habitat_visualization_data_gibson.py:13-36)."""
    import sys
    from oracle import qstep
    rng = np.random.default_rng(0)
    cells = [(5, 9), (2, 3), (7, 7)]
    for r, c in cells:
        for i in range(4):
            Image.fromarray(rng.integers(0, 256, (260, 340, 3), dtype=np.uint8)).save(
                os.path.join(tmp_path, f"{r}-{c}-{i}.jpg"), quality=92)
    sys.path.insert(0, REF)
    try:
        from dataloaders.habitat_visualization_data_gibson import HabitatQVisualizationDatasetGibson
    finally:
        sys.path.remove(REF)

    class Capture:
        def __init__(self, B):
            self.B, self.seen = B, []

        def __call__(self, frames):
            self.seen.append(frames.clone())
            return None, torch.zeros(self.B, 5), None

    for panorama in (False, True):
        cap = Capture(4)
        build_value_maps(str(tmp_path), cap, panorama=panorama, resolution=16, workers=2)
        assert len(cap.seen) == 4                                           # one (padded) batch per orientation
        order = list_cells(str(tmp_path))
        for ori in range(4):
            ds = HabitatQVisualizationDatasetGibson(str(tmp_path), orientation=ori, panorama=panorama)
            ref = {(r, c): im for r, c, im in (ds[k] for k in range(len(ds)))}
            assert sorted(ref) == order
            for b, cell in enumerate(order):
                mine = cap.seen[ori][b]
                mine = qstep.to_imgnet(mine if panorama else mine[None])    # [F,3,224,224]
                want = ref[cell] if panorama else ref[cell][None]
                assert (mine - want).abs().max().item() <= 1e-6, (panorama, ori, cell)
