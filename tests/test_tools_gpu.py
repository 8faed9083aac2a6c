"""GPU suite: the real-video harness end to end -- a directory of PNG frames through
tools/interpolate_dir.py (the reference's scripts/visualize_interpolation.py:105-199 loop): sliding window over the
directory, uint8 frames to the GPU, FullModel.interpolate_u8, output numbering, .flo intermediates."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n_frames,bottleneck,extra", [(2, "CONV", []), (4, "CLSTM", ["--save-flows", "--amp", "--channels-last"])])
def test_interpolate_dir_end_to_end(tmp_path, n_frames, bottleneck, extra):
    cv2 = pytest.importorskip("cv2")
    from ssm_b200 import formats
    src, dst = tmp_path / "in", tmp_path / "out"
    src.mkdir()
    rng = np.random.default_rng(0)
    base = cv2.GaussianBlur(rng.integers(0, 256, (90, 150, 3), dtype=np.uint8), (0, 0), 3)
    n_in, rate = 4, 4
    inputs = [np.roll(base, 2 * i, axis=1) for i in range(n_in)]
    for i, img in enumerate(inputs):
        cv2.imwrite(str(src / ("%03d.png" % i)), img)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "interpolate_dir.py"), "--input-dir", str(src), "--output-dir", str(dst),
           "--upsample-rate", str(rate), "--n-frames", str(n_frames), "--bottleneck", bottleneck] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    pngs = sorted(p for p in os.listdir(dst) if p.endswith(".png"))
    # visualize_interpolation.py numbering: first frame of every pair, its rate - 1 intermediate frames, ..., the last frame
    assert len(pngs) == (n_in - 1) * rate + 1
    assert pngs[0] == os.path.basename(formats.output_name(str(dst), 0))
    for k in range(n_in):                                        # the input frames sit at every rate-th slot, bit for bit
        got = cv2.imread(str(dst / pngs[k * rate]))
        assert np.array_equal(got, inputs[k]), "input frame %d is not reproduced at slot %d" % (k, k * rate)
    mid = cv2.imread(str(dst / pngs[rate // 2]))
    assert mid.shape == inputs[0].shape and mid.dtype == np.uint8
    if "--save-flows" in extra:
        flos = sorted(p for p in os.listdir(dst) if p.endswith(".flo"))
        assert len(flos) == 2 * (n_in - 1)
        f = formats.read_flo(str(dst / flos[0]))
        assert f.shape == (90, 150, 2) and np.isfinite(f).all()
