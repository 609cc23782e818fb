"""CPU checks around BASELINE configs[0] (batch 16, default torch init under seed 0, 1000 ancestral steps):

* ldt_b200.Score / ldt_b200.Compressor constructed after ``torch.manual_seed(0)`` hold the reference modules' random-init
  weights bit for bit (tests/golden/init_hashes.json, written by make_golden.py from the reference's own modules) -- so no
  weights are stored and the GPU test rebuilds them from the seed;
* the oracle reproduces the reference's own trajectory tensors (tests/golden/trajectory_b16.npz) teacher-forced;
* the golden recipe itself runs and regenerates committed fixtures bit-identically (build container only: it imports the
  unmodified reference from /root/reference).
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ldt_oracle as O
from tests.helpers import GOLDEN, airplane_config, golden, ns, rel_rms_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAJ_STEPS = (0, 1, 10, 100, 500, 998, 999)


def sd_hash(module) -> str:
    h = hashlib.sha256()
    for k, v in module.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def default_init_modules():
    """common_init(0) -> Score(cfg.score) -> Compressor(cfg.compressor): train_Latent_Diffusion.py:14-19."""
    from ldt_b200 import Compressor, Score
    c = ns(airplane_config())
    torch.manual_seed(0)
    return Score(c.score).eval(), Compressor(c.compressor).eval(), c


def test_default_init_is_the_reference_default_init_bit_for_bit():
    want = json.load(open(os.path.join(GOLDEN, "init_hashes.json")))
    score, comp, _ = default_init_modules()
    assert sd_hash(score) == want["score_sha256"]
    assert sd_hash(comp) == want["compressor_sha256"]


def test_oracle_reproduces_reference_trajectory_teacher_forced():
    """Steps 500 and 999 of the reference's own batch-16 run: params from the oracle net (4 samples, to keep the CPU suite
    short) within fp32 accumulation noise; x_mean / x_next of ALL golden steps from the oracle's Ancestral update given the
    reference's params and noise: bit-exact."""
    g = golden("trajectory_b16.npz")
    score, _, c = default_init_modules()
    sd = {k: v.detach() for k, v in score.state_dict().items()}
    sde = O.VPSDE(c.sde.beta_start, c.sde.beta_end, c.sde.sigma2_0, c.sde.sample_N)
    ts = torch.linspace(1.0, c.sde.sample_time_eps, c.sde.sample_N)
    with torch.no_grad():
        for i in (500, 999):
            t = torch.ones(4) * ts[i]
            params = O.score_forward(sd, c.score, g[f"x_{i}"][:4], t)
            assert rel_rms_err(params, g[f"params_{i}"][:4]) < 2e-5, (i, rel_rms_err(params, g[f"params_{i}"][:4]))
        for i in TRAJ_STEPS:
            t = torch.ones(16) * ts[i]
            x_next, x_mean = O.ancestral_step(sde, g[f"x_{i}"], t, g[f"params_{i}"], g[f"noise_{i}"], c.sde.sample_N)
            assert torch.equal(x_mean, g[f"xmean_{i}"]) and torch.equal(x_next, g[f"xnext_{i}"]), i
    assert torch.equal(g["xnext_0"], g["x_1"]) and torch.equal(g["xmean_999"], g["eps"])


def test_oracle_decoder_reproduces_reference_decode_of_unit_scale_latent():
    """Compressor.sample of the reference (default init, seed 0) on the loop's N(0,1) start latent, 2 clouds."""
    g = golden("trajectory_b16.npz")
    _, comp, c = default_init_modules()
    csd = {k: v.detach() for k, v in comp.state_dict().items()}
    with torch.no_grad():
        pts = O.decoder_sample(csd, c.compressor, g["x_0"][:2], 2048)
    assert rel_rms_err(pts, g["points_x0"][:2]) < 1e-4, rel_rms_err(pts, g["points_x0"][:2])


@pytest.mark.skipif(not os.path.isdir(os.environ.get("LDT_REFERENCE", "/root/reference")),
                    reason="needs the reference tree (build container only)")
def test_golden_recipe_runs_and_regenerates_committed_fixtures(tmp_path):
    """tests/golden/make_golden.py imports the UNMODIFIED reference; a repo-level package that shadows one of the
    reference's (tools/) would break it.  Regenerate nn / sde / decoder and compare bit for bit."""
    env = dict(os.environ, LDT_GOLDEN_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(GOLDEN, "make_golden.py"), "nn", "sde", "decoder"], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("nn.npz", "metrics.npz", "sde.npz", "sde_ext.npz", "decoder_full.npz"):
        with np.load(os.path.join(GOLDEN, name)) as a, np.load(os.path.join(str(tmp_path), name)) as b:
            assert sorted(a.files) == sorted(b.files), name
            for k in a.files:
                assert np.array_equal(a[k], b[k]), (name, k)
