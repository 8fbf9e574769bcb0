"""Value maps: the caller of the Q-network forward in `visualize_value.build_map_gibson`
(visualize_value.py:60-99; BASELINE configs[3]).

A value-map data folder holds pre-rendered views `{row}-{col}-{i}.jpg`, i = 0..3 (four headings per map
cell; dataloaders/habitat_visualization_data_gibson.py:13-36).  For each of four orientations the
reference pushes batches of 32 views -- one view (`images[orientation]`) or, for panorama networks, the
four views rotated by the orientation -- through `model(images).max(2).values` and scatters the [B, 5]
values into a `resolution x resolution x classes` map at (row, col), marking the cell in `free_map`.

`build_value_maps` does the same around a forward-only runner (`inference.QValueRunner`: uint8 frames in,
`max_a Q` out, one CUDA graph per call): views are decoded to uint8 in a thread pool (resize / centre
crop as `imageNetTransformPIL`; normalisation is the first kernel's job), the last batch is padded and
trimmed, and the scatter is the reference's `new_map[row, col] = values`.  Plotting (matplotlib,
:103-157) and the simulator-side rendering of the views are out of scope.
"""
from __future__ import annotations

import os
import re
from concurrent.futures import ThreadPoolExecutor
from typing import List, Tuple

import numpy as np
import torch

from . import realdata

BATCH = 32                                       # visualize_value.py:83
_NAME = re.compile(r"(\d+)-(\d+)-\d+\.jpg$")


def list_cells(data_folder: str) -> List[Tuple[int, int]]:
    """the (row, col) cells that have rendered views (habitat_visualization_data_gibson.py:18-22), sorted
    (the reference's order is a set's; the result does not depend on it)"""
    cells = set()
    for name in os.listdir(data_folder):
        m = _NAME.search(name)
        if m:
            cells.add((int(m[1]), int(m[2])))
    return sorted(cells)


def build_value_maps(data_folder: str, runner, *, panorama: bool, resolution: int = 1500, num_classes: int = 5,
                     workers: int = 6):
    """-> (maps: list of four [resolution, resolution, num_classes] float64 arrays, one per orientation;
    free_map [resolution, resolution]).  `runner(frames)` must return (Q, value [B, classes], best) for
    uint8 frames [B, 224, 224, 3] (or [B, 4, 224, 224, 3] when `panorama`), B = runner.B."""
    cells = list_cells(data_folder)
    B = runner.B
    realdata._ensure_resize()
    pin = torch.cuda.is_available()
    shape = (B, 4, 224, 224, 3)
    views = torch.empty(shape, dtype=torch.uint8, pin_memory=pin)       # all four headings of a batch of cells
    maps = [np.zeros((resolution, resolution, num_classes)) for _ in range(4)]
    free_map = np.zeros((resolution, resolution))
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        for lo in range(0, len(cells), B):
            chunk = cells[lo:lo + B]
            n = len(chunk)
            padded = chunk + [chunk[-1]] * (B - n)
            dst = views.numpy()
            jobs = [pool.submit(realdata.decode_frame, os.path.join(data_folder, f"{r}-{c}-{i}.jpg"), dst[b, i])
                    for b, (r, c) in enumerate(padded) for i in range(4)]
            for j in jobs:
                j.result()
            rows = np.array([r for r, _ in chunk])
            cols = np.array([c for _, c in chunk])
            for ori in range(4):                                           # :72
                if panorama:                                               # rotated_images = cat(images[ori:], images[:ori])
                    frames = torch.cat([views[:, ori:], views[:, :ori]], dim=1)
                else:                                                      # images[orientation]
                    frames = views[:, ori]
                _q, value, _best = runner(frames.contiguous())
                maps[ori][rows, cols] = value[:n].detach().cpu().double().numpy()     # new_map[row, col] = values (:97)
            free_map[rows, cols] = 1                                       # :98
    return maps, free_map
