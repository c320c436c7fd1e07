"""the ffmpeg-pipe frame feed/sink (host/videoio.hpp) and the reference-compatible command
line (host/main.cpp), exercised with fake `ffprobe` / `ffmpeg` shell scripts that speak raw bgr24 (the injectable command
prefix the reference lacks, writer.cpp:24).  CPU tests cover the pipes, the command lines and the argument rules
(main.cpp:17-153); the GPU test runs build + render end to end and compares every frame with Img2Img.render."""
import os
import stat
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "waifu2x-tensorrt_b200")
CLI = os.path.join(PKG, "bin", "waifu2x-b200")
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)

FAKE_FFPROBE = """#!/bin/bash
# prints the <input>.meta file the test wrote (key=value lines as `-of default=noprint_wrappers=1` would)
for last; do :; done
echo "$@" > "$last.probe_cmd"
cat "$last.meta"
"""

FAKE_FFMPEG = """#!/bin/bash
# reader (`-f image2pipe`): streams the input file, which already holds raw bgr24 frames;
# writer (`-i -`): copies stdin to the output path and records the command line next to it
args=("$@")
for last; do :; done
if [[ " $* " == *" image2pipe "* ]]; then
  for ((i = 0; i < ${#args[@]}; i++)); do
    if [[ "${args[$i]}" == "-i" ]]; then exec cat "${args[$((i + 1))]}"; fi
  done
else
  echo "$@" > "$last.cmd"
  exec cat > "$last"
fi
"""


def _fake_tools(d):
    for name, body in (("ffprobe", FAKE_FFPROBE), ("ffmpeg", FAKE_FFMPEG)):
        p = os.path.join(d, name)
        with open(p, "w") as f:
            f.write(body)
        os.chmod(p, os.stat(p).st_mode | stat.S_IXUSR)
    return d + "/"


def _fake_video(path, frames, rate="30000/1001", image=False):
    frames = np.asarray(frames, np.uint8)
    frames.tofile(path)
    n, h, w, _ = frames.shape
    with open(path + ".meta", "w") as f:
        f.write(f"width={w}\nheight={h}\nr_frame_rate={rate}\nnb_frames={'N/A' if image else n}\n")


VIDEOIO_PROG = r"""
#include "videoio.hpp"
#include <vector>
#include <cstdio>
int main(int argc, char** argv) {
    // argv: ffmpegDir input output
    try {
        VideoCapture cap;
        cap.setFfmpegDir(argv[1]);
        cap.open(argv[2]);
        const FrameSize sz = cap.getFrameSize();
        std::printf("%d %d %.6f %d\n", sz.width, sz.height, cap.getFrameRate(), cap.getFrameCount());
        VideoWriter wr;
        wr.setFfmpegDir(argv[1]).setFrameSize(sz).setFrameRate(cap.getFrameRate()).setOutputFile(argv[3]).setCodec("libx264")
            .setPixelFormat("yuv420p").setConstantRateFactor(23);
        wr.open();
        bool threw = false;
        try { wr.setCodec("x"); } catch (const std::exception&) { threw = true; }
        if (!threw) return 3;
        std::vector<uint8_t> buf(sz.bytes());
        int n = 0;
        while (cap.read(buf.data())) {
            for (auto& b : buf) b = 255 - b;
            wr.write(buf.data(), sz);
            ++n;
        }
        std::printf("%d %d\n", n, cap.getFrameIndex());
        try { wr.write(buf.data(), FrameSize{sz.width + 1, sz.height}); return 4; } catch (const std::invalid_argument&) {}
        cap.release();
        wr.release();
        try { cap.read(buf.data()); return 5; } catch (const std::runtime_error&) {}
        try { cap.open("/nonexistent/file.mp4"); return 6; } catch (const std::runtime_error&) {}
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
"""


def test_videoio_pipes_roundtrip(tmp_path):
    d = str(tmp_path)
    tools = _fake_tools(d)
    src = os.path.join(d, "prog.cpp")
    with open(src, "w") as f:
        f.write(VIDEOIO_PROG)
    exe = os.path.join(d, "prog")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(PKG, "host"), src, "-o", exe])
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (5, 12, 20, 3), dtype=np.uint8)
    vid = os.path.join(d, "clip.mp4")
    _fake_video(vid, frames)
    outp = os.path.join(d, "out.mp4")
    r = subprocess.run([exe, tools, vid, outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    l0, l1 = r.stdout.strip().splitlines()
    w, h, fps, n = l0.split()
    assert (int(w), int(h), int(n)) == (20, 12, 5) and abs(float(fps) - 30000 / 1001) < 1e-5
    assert l1.split() == ["5", "4"]  # frames read, last frame index (capture.cpp:126)
    got = np.fromfile(outp, np.uint8).reshape(frames.shape)
    assert np.array_equal(got, 255 - frames)
    # the command lines are the reference's (capture.cpp:65-68, writer.cpp:24-33)
    probe = open(vid + ".probe_cmd").read().strip()
    assert probe == ("-v error -select_streams v:0 -show_entries stream=width,height,r_frame_rate,nb_frames "
                     "-of default=noprint_wrappers=1 " + vid)
    cmd = open(outp + ".cmd").read().strip()
    assert cmd == f"-v error -y -f rawvideo -vcodec rawvideo -s 20x12 -pix_fmt bgr24 -r {30000 / 1001:.6f} -i - -vcodec libx264 -pix_fmt yuv420p -crf 23 {outp}"


def test_videoio_still_image_has_one_frame(tmp_path):
    """capture.cpp:88-92: ffprobe reports nb_frames=N/A for images; the capture then yields exactly one frame."""
    d = str(tmp_path)
    tools = _fake_tools(d)
    src = os.path.join(d, "prog.cpp")
    with open(src, "w") as f:
        f.write(VIDEOIO_PROG)
    exe = os.path.join(d, "prog")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(PKG, "host"), src, "-o", exe])
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (3, 8, 10, 3), dtype=np.uint8)   # the file holds more data than the one frame that is read
    img = os.path.join(d, "still.png")
    _fake_video(img, frames, rate="25/1", image=True)
    outp = os.path.join(d, "out.png")
    r = subprocess.run([exe, tools, img, outp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    l0, l1 = r.stdout.strip().splitlines()
    assert l0.split()[3] == "1" and abs(float(l0.split()[2]) - 25.0) < 1e-9
    assert l1.split() == ["1", "0"]
    assert np.array_equal(np.fromfile(outp, np.uint8).reshape(frames[0].shape), 255 - frames[0])


def test_videoio_invalid_probe(tmp_path):
    d = str(tmp_path)
    tools = _fake_tools(d)
    src = os.path.join(d, "prog.cpp")
    with open(src, "w") as f:
        f.write(VIDEOIO_PROG)
    exe = os.path.join(d, "prog")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(PKG, "host"), src, "-o", exe])
    vid = os.path.join(d, "bad.mp4")
    open(vid, "wb").write(b"xx")
    open(vid + ".meta", "w").write("bad.mp4: Invalid data found when processing input\n")
    r = subprocess.run([exe, tools, vid, os.path.join(d, "o.mp4")], capture_output=True, text=True)
    assert r.returncode == 1 and "input file is invalid" in r.stderr


def _cli(*args, cwd=None):
    return subprocess.run([CLI, *args], capture_output=True, text=True, cwd=cwd)


BASE = ["--model", "cunet/art", "--scale", "2", "--noise", "0", "--batchSize", "2", "--tileSize", "64"]


def test_cli_argument_rules(tmp_path, built_lib):  # built_lib: `make` builds bin/waifu2x-b200 next to lib/libw2x.so
    assert _cli("--help").returncode == 0
    assert "REQUIRED" in _cli("-h").stdout
    r = _cli("--model", "cunet/art", "build")
    assert r.returncode == 106 and "is required" in r.stderr
    assert _cli(*BASE).returncode == 106                                   # a subcommand is required (main.cpp:19)
    assert _cli(*BASE[:-1], "100", "build").returncode == 106             # tile size outside {64,128,256,400,640}
    assert _cli(*BASE, "--precision", "int8", "build").returncode == 106
    assert _cli(*BASE, "render").returncode == 106                         # -i required
    assert _cli(*BASE, "render", "-i", str(tmp_path / "nope.png")).returncode == 106
    assert _cli(*BASE, "render", "-i", str(tmp_path), "--blend", "0.3").returncode == 106
    assert _cli(*BASE, "render", "-i", str(tmp_path), "--crf", "52").returncode == 106
    assert _cli(*BASE, "build", "--tta").returncode == 106                 # render-only flag
    r = _cli("--model", "cunet/art", "--scale", "4", "--noise", "0", "--batchSize", "2", "--tileSize", "64", "build")
    assert r.returncode == 255 and "does not support scale factor 4" in r.stderr   # main.cpp:142-143
    r = _cli("--model", "swin_unet/art", "--scale", "1", "--noise", "-1", "--batchSize", "2", "--tileSize", "64", "build")
    assert r.returncode == 255 and "Noise level -1" in r.stderr                     # main.cpp:144-145


@pytest.mark.gpu
def test_cli_build_and_render_video_and_image(tmp_path, built_lib):
    import w2x
    from __graft_entry__ import make_synthetic_model

    d = str(tmp_path)
    tools = _fake_tools(d)
    models = os.path.join(d, "models")
    _, onnx_path = make_synthetic_model(models, 2, 0)
    r = _cli(*BASE, "build", cwd=d)                                        # models/ relative to the cwd (main.cpp:201)
    assert r.returncode == 0, r.stdout + r.stderr
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (7, 76, 100, 3), dtype=np.uint8)
    os.makedirs(os.path.join(d, "in", "sub"))
    os.makedirs(os.path.join(d, "out"))
    _fake_video(os.path.join(d, "in", "clip.mkv"), frames, rate="24/1")
    _fake_video(os.path.join(d, "in", "sub", "still.jpg"), frames[:1], rate="25/1", image=True)
    open(os.path.join(d, "in", "notes.txt"), "w").write("ignored")
    r = _cli(*BASE, "render", "-i", os.path.join(d, "in"), "--recursive", "-o", os.path.join(d, "out"), "--ffmpegDir", tools, cwd=d)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Rendered file" in r.stdout
    vid_out = os.path.join(d, "out", "clip(cunet_art)(noise0)(scale2).mp4")   # main.cpp:206-209,244-252
    img_out = os.path.join(d, "out", "still(cunet_art)(noise0)(scale2).png")
    assert os.path.exists(vid_out) and os.path.exists(img_out), os.listdir(os.path.join(d, "out"))

    eng = w2x.Img2Img()
    assert eng.load(onnx_path, w2x.RenderConfig(batchSize=2, height=64, width=64, scaling=2))
    want = np.stack([eng.render(f) for f in frames])
    eng.close()
    got = np.fromfile(vid_out, np.uint8).reshape(want.shape)
    assert np.array_equal(got, want)                                       # pipelined == serial, byte for byte
    assert np.array_equal(np.fromfile(img_out, np.uint8).reshape(want[0].shape), want[0])
    cmd = open(vid_out + ".cmd").read()
    assert "-s 200x152" in cmd and "-r 24.000000" in cmd and "-vcodec libx264 -pix_fmt yuv420p -crf 23" in cmd
    cmd = open(img_out + ".cmd").read()
    assert "-r 1.000000 -i - -crf 23" in cmd                               # image: codec/pix_fmt cleared (main.cpp:246-249)

    # extension: --device with a list shards the frames over an engine pool (the same GPU twice here) -- same bytes, same order
    os.makedirs(os.path.join(d, "out2"))
    r = _cli(*BASE, "--device", "0,0", "render", "-i", os.path.join(d, "in", "clip.mkv"), "-o", os.path.join(d, "out2"), "--ffmpegDir", tools, cwd=d)
    assert r.returncode == 0, r.stdout + r.stderr
    got2 = np.fromfile(os.path.join(d, "out2", "clip(cunet_art)(noise0)(scale2).mp4"), np.uint8).reshape(want.shape)
    assert np.array_equal(got2, want)
    # a file name with shell metacharacters is passed through intact (the reference would let the shell expand it)
    weird = os.path.join(d, "in", 'we$(ird)`x`".mkv')
    _fake_video(weird, frames[:2], rate="24/1")
    r = _cli(*BASE, "render", "-i", weird, "-o", os.path.join(d, "out2"), "--ffmpegDir", tools, cwd=d)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(os.path.join(d, "out2", 'we$(ird)`x`"(cunet_art)(noise0)(scale2).mp4'))

    # --nosuffix without -o writes next to the input with the new extension; missing engine => -1
    r = _cli(*BASE, "render", "-i", os.path.join(d, "in", "clip.mkv"), "--nosuffix", "--ffmpegDir", tools, "--tta", cwd=d)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.exists(os.path.join(d, "in", "clip.mp4"))
    r = _cli("--model", "cunet/art", "--scale", "2", "--noise", "3", "--batchSize", "2", "--tileSize", "64", "render", "-i",
             os.path.join(d, "in", "clip.mkv"), "--ffmpegDir", tools, cwd=d)
    assert r.returncode == 255
