"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference (Negai-98/LDT) on CPU.

Run in the build container only (``python tests/golden/make_golden.py``); /root/reference does not exist on
the GPU box, so the fixtures (small .npz files: inputs + reference outputs, never weights) are committed.
Weights are regenerated on both sides from ``oracle.ldt_oracle.synth_state_dict`` (seeded torch CPU generator),
loaded into the reference modules here and into ldt_b200 / the oracle in the tests.

The import shim is NOT a port: it only stubs modules absent from this image (torchdiffeq, pointnet2_ops,
mitsuba, matplotlib) and rewrites the reference's hard-coded ``device='cuda'`` strings to 'cpu'.
"""
import os
import sys
import types

import numpy as np
import torch
import yaml

REF = os.environ.get("LDT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "evaluation", "ChamferDistancePytorch"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


p2 = _stub("pointnet2_ops")
p2.pointnet2_utils = _stub("pointnet2_ops.pointnet2_utils", furthest_point_sample=lambda *a, **k: None)
_stub("torchdiffeq", odeint=None)
_stub("mitsuba")
mpl = _stub("matplotlib")
mpl.pyplot = _stub("matplotlib.pyplot")


class CudaToCpu(torch.overrides.TorchFunctionMode):
    """Rewrite device='cuda' / "cuda" arguments to CPU (reference hard-codes them, e.g. diffusion_continuous.py:638)."""

    def __torch_function__(self, func, types_, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if "device" in kwargs and str(kwargs["device"]).startswith("cuda"):
            kwargs["device"] = "cpu"
        args = tuple("cpu" if (isinstance(a, str) and a.startswith("cuda")) else a for a in args)
        return func(*args, **kwargs)


from oracle import ldt_oracle as O  # noqa: E402
from tests.helpers import airplane_config, small_score_cfg  # noqa: E402
from tools.io import dict2namespace  # noqa: E402  (reference)


OUT = os.environ.get("LDT_GOLDEN_OUT", HERE)   # tests/test_golden_recipe.py regenerates into a temp dir


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print(f"wrote {name}: " + ", ".join(f"{k}{tuple(np.asarray(v).shape)}" for k, v in arrays.items()),
          f"({os.path.getsize(path) / 1024:.1f} KiB)")


def load_synth(module, seed, gain=1.5):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = O.synth_state_dict(shapes, seed, gain)
    module.load_state_dict(sd, strict=True)
    return sd


def gen_score():
    from model.scorenet.score import Score
    for tag, cfg, seed, B in (("small", small_score_cfg(), 11, 3), ("full", dict2namespace(airplane_config()).score, 12, 2)):
        torch.manual_seed(0)
        model = Score(cfg).eval()
        load_synth(model, seed)
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn((B, cfg.z_scale, cfg.z_dim), generator=g)
        t = torch.rand((B,), generator=g) * 0.98 + 0.01
        with torch.no_grad():
            params = model(x, t)
            # per-block outputs for layer-wise debugging (first two blocks only, token-major)
            c = model.TimeEmbedding(t)
            h = model.ln_in(x.transpose(1, 2))
            h0 = model.Transformer[0](h, None, c)
        save(f"score_{tag}.npz", x=x, t=t, params=params, c=c, h_in=h.transpose(1, 2), h_block0=h0.transpose(1, 2))
        if tag == "small":
            # conditional variant: even blocks cross-attend to condition tokens, c += image vector (score.py:135,148)
            pts_cond = torch.randn((B, cfg.hidden_size, cfg.z_scale), generator=g)
            img_cond = torch.randn((B, cfg.t_dim), generator=g) * 0.5
            with torch.no_grad():
                params_c = model(x, t, condition=(pts_cond, img_cond))
            save("score_small_cond.npz", x=x, t=t, pts_cond=pts_cond, img_cond=img_cond, params=params_c)
        del model


def gen_unet():
    """The UNet wiring of the score net (score.py:67-83,138-146; default of model/scorenet/config.yaml)."""
    from model.scorenet.score import Score
    from tests.helpers import small_unet_score_cfg
    cfg = small_unet_score_cfg()
    torch.manual_seed(0)
    model = Score(cfg).eval()
    load_synth(model, 19)
    g = torch.Generator().manual_seed(119)
    x = torch.randn((3, cfg.z_scale, cfg.z_dim), generator=g)
    t = torch.rand((3,), generator=g) * 0.98 + 0.01
    img_cond = torch.randn((3, cfg.t_dim), generator=g) * 0.5
    with torch.no_grad():
        params = model(x, t)
        params_c = model(x, t, condition=(None, img_cond))
    save("score_unet.npz", x=x, t=t, img_cond=img_cond, params=params, params_cond=params_c)


def gen_condition():
    """ConditionNet + conditional Score from the reference, with the un-vendored pointnet2_ops FPS replaced by the C
    oracle (oracle/pointops_oracle.c); the same stand-in is used by the oracle-vs-golden test."""
    from tests.helpers import oracle_fps, small_cond_score_cfg
    sys.modules["pointnet2_ops.pointnet2_utils"].furthest_point_sample = lambda xyz, m: oracle_fps(xyz, m)
    from model.scorenet.score import Score
    cfg = small_cond_score_cfg()
    torch.manual_seed(0)
    with CudaToCpu():
        model = Score(cfg).eval()
        load_synth(model, 17)
        g = torch.Generator().manual_seed(117)
        B = 3
        img = torch.rand((B, 3, 64, 64), generator=g)
        pts = torch.randn((B, 600, 3), generator=g)
        pts = pts / pts.norm(dim=-1).max(dim=1)[0][:, None, None]
        x = torch.randn((B, cfg.z_scale, cfg.z_dim), generator=g)
        t = torch.rand((B,), generator=g) * 0.98 + 0.01
        with torch.no_grad():
            pts_cond, img_cond = model.c_net({"img": img, "pts": pts})
            params = model(x, t, condition={"img": img, "pts": pts})
            params_pts_only = model(x, t, condition={"pts": pts})
    save("condition.npz", img=img, pts=pts, x=x, t=t, pts_cond=pts_cond, img_cond=img_cond, params=params,
         params_pts_only=params_pts_only)


def gen_encoder():
    """Compressor.forward (bottom_up + top_down, Network.py:188-249) of the reference in eval mode; the un-vendored
    pointnet2_ops FPS is replaced by the C oracle as in gen_condition()."""
    from tests.helpers import oracle_fps
    sys.modules["pointnet2_ops.pointnet2_utils"].furthest_point_sample = lambda xyz, m: oracle_fps(xyz, m)
    from model.Compressor.Network import Compressor
    cfg = dict2namespace(airplane_config()).compressor
    torch.manual_seed(0)
    comp = Compressor(cfg).eval()
    load_synth(comp, 13, gain=0.6)   # 18 chained blocks with sharp 2048-key softmaxes: keep the map well conditioned
    g = torch.Generator().manual_seed(213)
    pts = torch.randn((2, 2048, 3), generator=g)
    pts = pts / pts.norm(dim=-1).max(dim=1)[0][:, None, None]
    with CudaToCpu(), torch.no_grad():
        torch.manual_seed(6)
        out = comp(pts)
    save("encoder.npz", pts=pts, set=out["set"], all_eps=out["all_eps"], kls=torch.stack(out["kls"]),
         mu=torch.stack([p[1] for p in out["posteriors"][1:]]), logvar=torch.stack([p[2] for p in out["posteriors"][1:]]),
         max=out["max"])


def gen_decoder():
    from model.Compressor.Network import Compressor
    cfg = dict2namespace(airplane_config()).compressor
    torch.manual_seed(0)
    comp = Compressor(cfg).eval()
    load_synth(comp, 13)
    g = torch.Generator().manual_seed(113)
    eps = torch.randn((2, cfg.z_scales, cfg.n_layers * cfg.z_dim), generator=g)
    with CudaToCpu(), torch.no_grad():
        torch.manual_seed(5)
        full = comp.sample((2, 2048), given_eps=eps)
        torch.manual_seed(5)
        part = comp.sample((2, 1000), given_eps=eps)
    save("decoder_full.npz", eps=eps, points_2048=full, points_1000=part)


def gen_sde():
    with CudaToCpu():
        from diffusion.diffusion_continuous import DiffusionVPSDE
        cfg = dict2namespace(airplane_config()).sde
        sde = DiffusionVPSDE(cfg)
        ts = torch.linspace(1.0, 1e-6, 1000)
        out = dict(betas=sde.betas, alphas_cump=sde.alphas_cump, timesteps=ts, var=sde.var(ts), g2=sde.g2(ts),
                   e2int_f=sde.e2int_f(ts))
        # drive the reference sampler itself for a few steps with a deterministic stand-in score network and
        # record the noise it draws, once per predictor
        real_randn_like = torch.randn_like
        for pred in ("ancestral", "reversediffusion", "eulermaruyama", "ddim"):
            for denoise in (True, False):
                noises = []

                def rec(x, *a, **k):
                    z = real_randn_like(x, *a, **k)
                    noises.append(z.clone())
                    return z

                def score_fn(t, x, label=None, condition=None):
                    params = 0.3 * x + torch.sin(5.0 * t)[:, None, None]
                    return -params / torch.sqrt(sde.var(t))[:, None, None], params

                torch.manual_seed(21)
                torch.randn_like = rec
                try:
                    res = sde.sample_discrete(score_fn=score_fn, num_samples=2, N=6, predictor=pred, corrector=None,
                                              corrector_steps=1, shape=(32, 120), time_eps=1e-6, probability_flow=False,
                                              denoise=denoise, snr=0.01, device="cpu")
                finally:
                    torch.randn_like = real_randn_like
                torch.manual_seed(21)
                x0 = torch.randn((2, 32, 120))
                key = f"{pred}_{'mean' if denoise else 'x'}"
                out[key] = res
                if denoise:
                    out[f"{pred}_x0"] = x0
                    out[f"{pred}_noise"] = torch.stack(noises)
        save("sde.npz", **out)

        # correctors, print_steps trajectory and PNDM.  The reference's `step_size[:, None]` /
        # `at.view(-1, 1)` broadcasts only type-check when the batch equals shape[0] (or 1): batch 32, shape (32, 4).
        ext = {}

        def run(tag, sde_obj=sde, **kw):
            noises = []

            def rec(x, *a, **k):
                z = real_randn_like(x, *a, **k)
                noises.append(z.clone())
                return z

            def score_fn(t, x, label=None, condition=None):
                params = 0.3 * x + torch.sin(5.0 * t)[:, None, None]
                return -params / torch.sqrt(sde_obj.var(t))[:, None, None], params

            args = dict(score_fn=score_fn, num_samples=32, N=5, predictor="ancestral", corrector=None, corrector_steps=1,
                        shape=(32, 4), time_eps=1e-6, probability_flow=False, denoise=True, snr=0.16, device="cpu")
            args.update(kw)
            torch.manual_seed(33)
            torch.randn_like = rec
            try:
                res = sde_obj.sample_discrete(**args)
            finally:
                torch.randn_like = real_randn_like
            torch.manual_seed(33)
            ext[f"{tag}_x0"] = torch.randn((32, 32, 4))
            ext[f"{tag}_out"] = torch.stack(res) if isinstance(res, list) else res
            if noises:
                ext[f"{tag}_noise"] = torch.stack(noises)

        run("pc_ancestral", corrector="ancestral", corrector_steps=2)
        run("pc_langevin", corrector="langevin", corrector_steps=1, predictor="eulermaruyama")
        run("c_only_ancestral", predictor=None, corrector="ancestral", corrector_steps=1, denoise=False)
        run("print_steps", print_steps=4)
        cfg2 = dict2namespace(airplane_config()).sde
        cfg2.sample_N = 6
        run("pndm", sde_obj=DiffusionVPSDE(cfg2), predictor="pndm")
        save("sde_ext.npz", **ext)


def _sd_hash(module):
    import hashlib
    h = hashlib.sha256()
    for k, v in module.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


TRAJ_STEPS = (0, 1, 10, 100, 500, 998, 999)


def gen_trajectory():
    """BASELINE configs[0] run by the reference itself on CPU: common_init(0) -> Score(cfg.score), Compressor(cfg.compressor)
    with torch's DEFAULT init (initialize_weights is never called, score.py:98), batch 16, the full 1000-step ancestral
    loop (diffusion_continuous.py:152-162, 242-249) through the Trainer.score_fn closure (trainer/Latent_SDE_Trainer.py:
    57-61), then Compressor.sample.  Stores, for steps TRAJ_STEPS, the loop state x_i, the noise drawn, params, x_mean and
    x_next (SURVEY.md 8d teacher-forced parity inputs), the final latent and the decoded points -- never weights: both
    sides rebuild them from the seed (init_hashes.json pins that they are the same bits).  ~15 min on 8 cores."""
    import json
    from diffusion.diffusion_continuous import DiffusionVPSDE
    from model.Compressor.Network import Compressor
    from model.scorenet.score import Score
    cfg = dict2namespace(airplane_config())
    B = 16
    with CudaToCpu():
        torch.manual_seed(0)                       # tools/utils.py:269-276 common_init(seed=0)
        model = Score(cfg.score).eval()            # train_Latent_Diffusion.py:18-19 construction order
        comp = Compressor(cfg.compressor).eval()
        with open(os.path.join(OUT, "init_hashes.json"), "w") as f:
            json.dump({"seed": 0, "score_sha256": _sd_hash(model), "compressor_sha256": _sd_hash(comp),
                       "how": "sha256 over (key, fp32 bytes) of state_dict() after torch.manual_seed(0); Score(cfg.score); "
                              "Compressor(cfg.compressor) with the reference modules"}, f, indent=1)
        sde = DiffusionVPSDE(cfg.sde)
        rec, call = {}, [0]
        reuse = os.path.join(HERE, "trajectory_b16.npz")
        if os.environ.get("LDT_TRAJ_REUSE") == "1" and os.path.exists(reuse):
            # add / refresh the decoder outputs without repeating the 15-minute loop: the stored loop tensors are kept
            with np.load(reuse) as z:
                old = {k: torch.from_numpy(z[k]) for k in z.files}
            with torch.no_grad():
                torch.manual_seed(99)
                old["points"] = comp.sample((B, 2048), given_eps=old["eps"])
                old["points_x0"] = comp.sample((B, 2048), given_eps=old["x_0"])
            save("trajectory_b16.npz", **old)
            return

        def score_fn(t, x, label=None, condition=None):   # trainer/Latent_SDE_Trainer.py:57-61
            t = t.to(x)
            params = model(x, t, label=label, condition=condition)
            var = sde.var(t)[:, None, None]
            i = call[0]
            if i in TRAJ_STEPS:
                rec[f"x_{i}"], rec[f"params_{i}"] = x.clone(), params.clone()
            if i - 1 in TRAJ_STEPS:
                rec[f"xnext_{i - 1}"] = x.clone()
            call[0] += 1
            if i % 50 == 0:
                print("  step", i, flush=True)
            return -params / torch.sqrt(var), params

        real_randn_like = torch.randn_like

        def rec_randn_like(x, *a, **k):
            z = real_randn_like(x, *a, **k)
            i = call[0] - 1
            if i in TRAJ_STEPS:
                rec[f"noise_{i}"] = z.clone()
                rec[f"xmean_{i}"] = sys._getframe(1).f_locals["x_mean"].clone()   # Ancestral's local (:159)
            return z

        torch.manual_seed(1234)                    # sampling seed
        torch.randn_like = rec_randn_like
        try:
            with torch.no_grad():
                eps = sde.sample_discrete(score_fn=score_fn, num_samples=B, N=cfg.sde.sample_N, predictor="ancestral",
                                          corrector=None, corrector_steps=1, shape=(cfg.score.z_scale, cfg.score.z_dim),
                                          time_eps=cfg.sde.sample_time_eps, probability_flow=False, denoise=True,
                                          snr=cfg.sde.snr, device="cpu")
        finally:
            torch.randn_like = real_randn_like
        last = TRAJ_STEPS[-1]
        rec[f"xnext_{last}"] = rec[f"xmean_{last}"] + torch.sqrt(sde.betas[0]) * rec[f"noise_{last}"]   # :161 (idx 0)
        assert torch.equal(eps, rec[f"xmean_{last}"])
        with torch.no_grad():
            pts = comp.sample((B, 2048), given_eps=eps)   # trainer/Latent_SDE_Trainer.py:163 (CPU randperms follow on)
            # the same decoder on a unit-scale latent (the loop's N(0,1) start): with random-init nets the FINAL latent has
            # rms ~280, which saturates the decoder's softmaxes and makes that decode ill-conditioned in any arithmetic
            pts0 = comp.sample((B, 2048), given_eps=rec["x_0"])
    save("trajectory_b16.npz", eps=eps, points=pts, points_x0=pts0, **rec)


def gen_layout():
    """state_dict key -> shape of the reference modules for the shipped config (the checkpoint contract)."""
    import json
    from model.Compressor.Network import Compressor
    from model.scorenet.score import Score
    cfg = dict2namespace(airplane_config())
    cfg.score.num_blocks = 2  # layout per block is identical; keeps the instantiation light
    from tests.helpers import small_cond_score_cfg, small_unet_score_cfg
    lay = {"score_unet_small": [[k, list(v.shape)] for k, v in Score(small_unet_score_cfg()).state_dict().items()],
           "score_cond_small": [[k, list(v.shape)] for k, v in Score(small_cond_score_cfg()).state_dict().items()],
           "score_2blocks": [[k, list(v.shape)] for k, v in Score(cfg.score).state_dict().items()],
           "compressor": [[k, list(v.shape)] for k, v in Compressor(cfg.compressor).state_dict().items()]}
    with open(os.path.join(HERE, "state_dict_layout.json"), "w") as f:
        json.dump(lay, f)
    print("wrote state_dict_layout.json:", {k: len(v) for k, v in lay.items()})


def gen_nn():
    import chamfer_python  # reference evaluation/ChamferDistancePytorch/chamfer_python.py (float64 brute force)
    g = torch.Generator().manual_seed(31)
    p1 = torch.rand((4, 100, 3), generator=g)
    p2 = torch.rand((4, 200, 3), generator=g)
    d1, d2, i1, i2 = chamfer_python.distChamfer(p1, p2)  # the check used by unit_test.py:22-33
    # ragged / tie cases: duplicated points => ties must resolve to the lowest index
    q1 = torch.rand((2, 37, 3), generator=g)
    q2 = torch.cat([q1[:, :5], torch.rand((2, 60, 3), generator=g), q1[:, :5]], dim=1)
    e1, e2, j1, j2 = chamfer_python.distChamfer(q1, q2)
    save("nn.npz", p1=p1, p2=p2, dist1=d1, dist2=d2, idx1=i1, idx2=i2, q1=q1, q2=q2, qdist1=e1, qdist2=e2)

    from evaluation import evaluation_metrics as EM  # reference metrics (falls back to bmm distChamfer on CPU)
    ref = torch.rand((6, 64, 3), generator=g)
    smp = torch.rand((5, 64, 3), generator=g) * 0.9
    M_rs = EM._pairwise_CD_(ref, smp, 4)
    M_rr = EM._pairwise_CD_(ref, ref, 4)
    M_ss = EM._pairwise_CD_(smp, smp, 4)
    mc = EM.lgan_mmd_cov(M_rs.t())
    kn = EM.knn(M_rr, M_rs, M_ss, 1, sqrt=False)
    save("metrics.npz", ref=ref, smp=smp, M_rs=M_rs, M_rr=M_rr, M_ss=M_ss, mmd=mc["mmd"], cov=mc["cov"],
         acc=kn["acc"], tp=kn["tp"], fp=kn["fp"], fn=kn["fn"], tn=kn["tn"])


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    which = sys.argv[1:] or ["score", "unet", "condition", "encoder", "decoder", "sde", "nn", "layout"]
    for w in which:
        {"score": gen_score, "unet": gen_unet, "condition": gen_condition, "encoder": gen_encoder, "decoder": gen_decoder, "sde": gen_sde, "nn": gen_nn, "layout": gen_layout, "trajectory": gen_trajectory}[w]()
