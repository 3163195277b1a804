"""Host-side dataset loop of the path: object_detect::findObjectDataset (reference
src/libs/libPictStruct/objectdetect_aux.cpp:322-406) and the image-range sharding of the CLI
(src/apps/partapp/main.cpp:155-192: --first/--numimgs, --distribute/--ncpu/--batch_num).

Images are independent, so ranks (GPUs) take contiguous image-index ranges and results are gathered on the host --
no collective touches the data path.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def init_firstidx_lastidx(num_images: int, first: Optional[int] = None, numimgs: Optional[int] = None) -> Tuple[int, int]:
    """main.cpp:155-192 init_firstidx_lastidx: [first, last] inclusive, clipped to the list."""
    firstidx = 0 if first is None else int(first)
    if firstidx < 0 or firstidx > num_images:
        raise ValueError("first image index out of range (aux.cpp:328 assert)")
    lastidx = num_images - 1 if numimgs is None else min(num_images - 1, firstidx + int(numimgs) - 1)
    return firstidx, lastidx


def shard_range(firstidx: int, lastidx: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous sub-range [lo, hi] (inclusive; hi < lo means empty) of [firstidx, lastidx] for `rank` of `world`,
    the same chunking --distribute --ncpu N --batch_num k produces (main.cpp:267-394): ceil(n/world) images per batch."""
    n = lastidx - firstidx + 1
    if n <= 0:
        return firstidx, firstidx - 1
    per = -(-n // world)
    lo = firstidx + rank * per
    hi = min(lastidx, lo + per - 1)
    return lo, hi


def get_object_hyp_filename(imgidx: int, flip: bool, spm: str = "none") -> str:
    """aux.cpp:311-320 getObjectHypFilename"""
    return "/object_hyp_imgidx%d_o%d_spm%s.pbuf" % (imgidx, int(flip), spm)


def pose_est_filename(imgidx: int) -> str:
    """findrot.cpp:1010: pose_est_imgidx%04d.mat"""
    return "/pose_est_imgidx%04d.mat" % imgidx


def find_object_dataset(infer_image: Callable[[int, bool], np.ndarray], firstidx: int, lastidx: int,
                        flip_orientation: bool = False, rank: int = 0, world: int = 1,
                        gather: Optional[Callable[[np.ndarray], List[np.ndarray]]] = None):
    """findObjectDataset (aux.cpp:368-402): image x flip loop over this rank's shard.  `infer_image(imgidx, flip)`
    returns the best_conf rows [P][7] of one image.  Returns {(imgidx, flip): best_conf} for this rank, or for all
    ranks when `gather` (e.g. an all_gather of a padded array) is given."""
    lo, hi = shard_range(firstidx, lastidx, rank, world)
    flips = (False, True) if flip_orientation else (False,)
    local = {}
    for imgidx in range(lo, hi + 1):
        for flip in flips:
            local[(imgidx, flip)] = np.asarray(infer_image(imgidx, flip), np.float32)
    if gather is None:
        return local
    # pack -> gather -> unpack; keys travel as two extra leading floats so the ranks need no shared state
    rows = [np.concatenate([[k[0], float(k[1])], v.reshape(-1)]) for k, v in sorted(local.items())]
    width = max((len(r) for r in rows), default=0)
    per = -(-(lastidx - firstidx + 1) // world) * len(flips)
    packed = np.full((per, max(width, 2)), np.nan, np.float32)
    for i, r in enumerate(rows):
        packed[i, :len(r)] = r
    out = {}
    for part in gather(packed):
        for r in part:
            if not np.isnan(r[0]):
                out[(int(r[0]), bool(r[1]))] = r[2:].reshape(-1, 7)
    return out
