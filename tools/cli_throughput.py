"""Times the C++ drop-in host (`psinfer_partapp --find_obj`) on a synthetic experiment written to disk in the reference's
formats: images/s at 1 GPU (and at every GPU count the box offers), and a byte comparison of the outputs between runs.

    python tools/cli_throughput.py [--images 256] [--distinct 16] [--out gpurun_out/cli_throughput.json]

The score-grid files of image i >= --distinct are symlinks to those of image i % --distinct (the host still opens,
inflates and uploads every one of them); everything else is per image.
"""
import argparse
import filecmp
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.make_experiment import make  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--contexts", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cli_throughput.json"))
    args = ap.parse_args()
    import torch
    ngpu = torch.cuda.device_count()
    subprocess.run(["make", "-C", os.path.join(ROOT, "partapp_b200", "csrc", "host")], check=True, capture_output=True)
    cli = os.path.join(ROOT, "partapp_b200", "psinfer_partapp")
    tmp = tempfile.mkdtemp(prefix="psinfer_cli_")
    t0 = time.time()
    info = make(os.path.join(tmp, "exp"), num_images=args.images, P=10, R=24, H=600, W=400, link_after=args.distinct)
    gen_s = time.time() - t0
    base = info["base"]
    runs = []
    ref_dir = None
    for gpus in sorted({1, 2, 4, 8} & set(range(1, ngpu + 1))):
        for sub in ("part_marginals", "object_hyp"):
            shutil.rmtree(os.path.join(base, sub), ignore_errors=True)
        t0 = time.time()
        r = subprocess.run([cli, "--expopt", info["expopt"], "--find_obj", "--gpus", str(gpus), "--contexts", str(args.contexts)],
                           capture_output=True, text=True, env=dict(os.environ, PSINFER_HOST_TIMING="1"))
        wall = time.time() - t0
        assert r.returncode == 0, r.stderr
        m = re.search(r"in ([0-9.]+) s \(([0-9.]+) images/s\)", r.stdout)
        run = {"gpus": gpus, "contexts_per_gpu": args.contexts, "images": args.images, "wall_s": round(wall, 2),
               "loop_s": float(m.group(1)), "images_per_s": float(m.group(2)), "host_cpus": os.cpu_count(),
               "host_phases": r.stderr.strip().splitlines()[-1] if r.stderr.strip() else None}
        keep = os.path.join(tmp, "out_%d" % gpus)
        os.makedirs(keep)
        for sub in ("part_marginals", "object_hyp"):
            shutil.copytree(os.path.join(base, sub), os.path.join(keep, sub))
        if ref_dir is None:
            ref_dir = keep
            run["identical_to_1gpu"] = True
        else:
            same = True
            for sub in ("part_marginals", "object_hyp"):
                names = sorted(os.listdir(os.path.join(ref_dir, sub)))
                same = same and names == sorted(os.listdir(os.path.join(keep, sub)))
                same = same and all(filecmp.cmp(os.path.join(ref_dir, sub, n), os.path.join(keep, sub, n), shallow=False) for n in names)
            run["identical_to_1gpu"] = bool(same)
        runs.append(run)
        print(json.dumps(run))
    out = {"workload": "configs[1]/[2] on disk: %d images, 10 parts, R=24, 600x400 (score grids: %d distinct files per part, the rest "
                       "symlinks), written with scipy.io in the reference's formats" % (args.images, args.distinct),
           "generation_s": round(gen_s, 1), "runs": runs}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
