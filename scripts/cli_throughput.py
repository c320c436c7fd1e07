#!/usr/bin/env python
"""End-to-end frames/s of the command line driver on a 1080p clip with stand-in ffprobe/ffmpeg scripts that move raw bgr24
(no codec in the image): decode pipe -> reader thread -> pinned ring -> GPU -> writer thread -> encode pipe.
Usage (GPU box): python scripts/cli_throughput.py [frames=96]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "waifu2x-tensorrt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_cli_videoio import CLI, _fake_tools, _fake_video  # noqa: E402
from __graft_entry__ import make_synthetic_model  # noqa: E402
from oracle import tiling  # noqa: E402

frames_n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    tools = _fake_tools(d)
    make_synthetic_model(os.path.join(d, "models"), 2, 3)
    base = ["--model", "cunet/art", "--scale", "2", "--noise", "3", "--batchSize", "8", "--tileSize", "256"]
    subprocess.check_call([CLI, *base, "build"], cwd=d, stdout=subprocess.DEVNULL)
    frames = np.stack([tiling.synthetic_frame(1920, 1080, s) for s in range(4)])
    clip = os.path.join(d, "clip.mkv")
    _fake_video(clip, np.concatenate([frames] * (frames_n // 4)), rate="24/1")
    # the sink discards the frames (an encoder would consume them); the command line is still recorded
    with open(os.path.join(d, "ffmpeg"), "w") as f:
        f.write('#!/bin/bash\nargs=("$@")\nif [[ " $* " == *" image2pipe "* ]]; then for ((i=0;i<${#args[@]};i++)); do '
                'if [[ "${args[$i]}" == "-i" ]]; then exec cat "${args[$((i+1))]}"; fi; done; else exec cat > /dev/null; fi\n')
    t0 = time.perf_counter()
    r = subprocess.run([CLI, *base, "render", "-i", clip, "--ffmpegDir", tools], cwd=d, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    n = (frames_n // 4) * 4
    print(json.dumps({"what": "waifu2x-b200 render, 1920x1080 -> 3840x2160, cunet/art 2x noise3 tile256 batch8, raw bgr24 pipes",
                      "frames": n, "wall_s": dt, "fps_incl_startup": n / dt, "output_mpx_s_incl_startup": n * 8.2944 / dt,
                      "note": "wall clock of the whole process: CUDA init, engine load and pipe startup included"}))
