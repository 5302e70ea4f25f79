"""GPU coverage of the step's tail and of the callers around it:

* ``csrc/optim.cu`` (``jpb_sumsq`` + ``jpb_adam_step``) against ``torch.optim.Adam`` + ``clip_grad_norm_`` on cuda:0;
* the NCCL gradient exchange on two GPUs: exchanged gradient == mean of the per-rank gradients, parameters stay identical
  (reference semantics: mono/core/utils/dist_utils.py:34-60) — skipped on a one-GPU box, run with ``gpurun --gpus 2``;
* ``train_mono`` + ``Runner`` + checkpoint + resume + ``DistEvalMonoHook`` on cuda with the real ``Baseline`` and the synthetic
  dataset, driven by a config file in the reference's format (what the reference's ``train.py`` does, trainer.py:146-199).
"""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from jperceiver_b200 import _lib  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def real_library():
    _lib._handle, _lib._emulated = None, False
    _lib.lib()
    yield


class _Holder(torch.nn.Module):
    def __init__(self, shapes, dev, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes])


@pytest.mark.parametrize("max_norm,world", [(None, 1), (35.0, 1), (0.5, 4)])
def test_adam_and_sumsq_kernels_match_torch(max_norm, world):
    """Flat clip + Adam (one ``jpb_sumsq`` + one ``jpb_adam_step`` per step, 1/world scaling fused) vs the reference chain:
    grad /= world -> clip_grad_norm_ -> Adam.step, six steps on 1.3 M parameters of ragged shapes."""
    from jperceiver_b200.apis.trainer import FlatParameters, FusedAdam
    dev = torch.device("cuda:0")
    shapes = [(256, 128, 3, 3), (513,), (64, 7, 7, 3), (1000, 511), (1,), (33, 65, 5)]
    a, b = _Holder(shapes, dev, 1), _Holder(shapes, dev, 1)
    ref = torch.optim.Adam(a.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    flat = FlatParameters(b)
    opt = FusedAdam(flat, lr=1e-3, max_norm=max_norm)
    g = torch.Generator().manual_seed(7)
    for it in range(6):
        scale = 10.0 ** (it - 3)                        # gradient norms below and above the clip threshold
        flat.zero_grad()
        for pa, pb in zip(a.parameters(), b.parameters()):
            gr = (torch.randn(pa.shape, generator=g) * scale).to(dev)
            pa.grad = gr.clone() / world                # reference: allreduce_grads divides by the world size first
            pb.grad.copy_(gr)                           # ours: the summed gradient; 1/world is applied inside the kernel
        if max_norm:
            total = torch.nn.utils.clip_grad_norm_(list(a.parameters()), max_norm)
        ref.step()
        opt.step(world)
        if max_norm:
            assert abs(opt.grad_norm(world) - float(total)) <= 1e-5 * float(total)
    torch.cuda.synchronize()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert (pa - pb).abs().max().item() <= 2e-6 + 1e-5 * pa.abs().max().item()
    assert int(opt.step_count.item()) == 6


def _nccl_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from jperceiver_b200.apis import TrainEngine
    from jperceiver_b200.model import MONO
    from jperceiver_b200 import synthetic
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    opt = dict(name="Baseline", depth_num_layers=18, pose_num_layers=18, frame_ids=[0, -1], imgs_per_gpu=1, height=128, width=384,
               scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0, depth_pretrained_path=None, pose_pretrained_path=None,
               automask=True, disp_norm=True, smoothness_weight=1e-3, scale_weight=0.1, dynamic_weight=15.0, static_weight=5.0,
               occ_map_size=64, num_class=2, loss_type="iou", loss_weight=20, loss2_type="boundary", loss2_weight=20,
               type="static_eigen", loss_sum=3, split="odometry")
    os.environ["JPB_OVERLAP_ALLREDUCE"] = "1"           # exercise the bucketed exchange (opt-in: measured slower than one piece)
    torch.manual_seed(1234 + 17 * rank)                 # DIFFERENT initial weights per rank: the engine must broadcast rank 0's
    model = MONO.module_dict["Baseline"](opt).to(dev).train()
    eng = TrainEngine(model, dict(type="Adam", lr=1e-4, weight_decay=0), dict(max_norm=35, norm_type=2))
    p0 = eng.flat.param.clone()
    allp = [torch.zeros_like(p0) for _ in range(world)]
    dist.all_gather(allp, p0)
    ok_bcast = all(torch.equal(allp[0], p) for p in allp)
    bufs = torch.cat([b.detach().float().reshape(-1) for b in model.buffers()])
    allb = [torch.zeros_like(bufs) for _ in range(world)]
    dist.all_gather(allb, bufs)
    ok_bcast = ok_bcast and all(torch.equal(allb[0], b) for b in allb)
    host = synthetic.make_batch(opt, 1, seed=100 + rank)
    from jperceiver_b200.apis import change_input_variable
    data = change_input_variable(host, dev)
    eng.forward_backward(data)                          # per-rank gradient of this rank's shard (BN statistics stay per rank)
    local = eng.flat.grad.clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    expect = sum(gathered) / world
    eng.exchange_gradients()
    got = eng.flat.grad / world
    den = expect.abs().max().item()
    ok_grad = (got - expect).abs().max().item() <= 1e-6 * max(den, 1e-12)
    eng.optimizer.step(world)
    for _ in range(3):                                  # full steps through the public entry: the first traces the completion order,
        eng.step(data)                                  # the next ones all-reduce bucket by bucket during backward
    assert eng.exchange is not None and eng.exchange.mode == "run" and len(eng.exchange.buckets) >= 4
    params = eng.flat.param.clone()
    allp = [torch.zeros_like(params) for _ in range(world)]
    dist.all_gather(allp, params)
    ok_sync = all(torch.equal(allp[0], p) for p in allp)
    finite = bool(torch.isfinite(params).all().item())
    q.put((rank, bool(ok_bcast), bool(ok_grad), bool(ok_sync), finite))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_gradient_exchange_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2); the same logic runs on gloo in tests/test_trainer.py")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(120)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(all(r[1:]) for r in res), res


CFG = '''
# a config file in the reference's format (config/cfg_kitti_baseline_odometry_boundary_ce_iou_1024_20.py), reduced shapes
HEIGHT, WIDTH, IMGS_PER_GPU = 128, 384, 2
data = dict(name='synthetic', split='odometry', height=HEIGHT, width=WIDTH, frame_ids=[0, -1, 1], num_samples=8, occ_map_size=64,
            in_path='', gt_depth_path='', png=False, stereo_scale=False)
model = dict(name='Baseline', depth_num_layers=18, pose_num_layers=18, frame_ids=[0, -1, 1], imgs_per_gpu=IMGS_PER_GPU, height=HEIGHT,
             width=WIDTH, scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0, depth_pretrained_path=None, pose_pretrained_path=None,
             automask=True, disp_norm=True, smoothness_weight=1e-3, scale_weight=0.1, dynamic_weight=15.0, static_weight=5.0,
             occ_map_size=64, num_class=2, loss_type='iou', loss_weight=20, loss2_type='boundary', loss2_weight=20, type='static',
             loss_sum=3, split='odometry')
resume_from = None
load_from = None
imgs_per_gpu = IMGS_PER_GPU
workers_per_gpu = 0
validate = True
validate_interval = 1
optimizer = dict(type='Adam', lr=1e-4, weight_decay=0)
optimizer_config = dict(grad_clip=dict(max_norm=35, norm_type=2))
lr_config = dict(policy='step', warmup=None, step=[1], gamma=0.5)
checkpoint_config = dict(interval=1)
log_config = dict(interval=1, hooks=[dict(type='TextLoggerHook')])
total_epochs = 2
dist_params = dict(backend='nccl')
log_level = 'INFO'
workflow = [('train', 1)]
gpus = [0]
'''


def test_train_mono_runner_checkpoint_resume_eval_on_cuda(tmp_path):
    """What the reference's ``train.py`` does after parsing its arguments (train.py:51-99 -> trainer.py:58-73,146-199), on cuda:0:
    config file -> datasets -> ``MONO.module_dict[name](cfg.model)`` -> ``train_mono(..., validate=True)`` for two epochs with the
    step LR policy, per-epoch checkpoints in the reference's format, JSON log lines and the device-side ``DistEvalMonoHook``;
    then a fresh process-equivalent resumes from ``latest.pth`` and continues."""
    from jperceiver_b200.apis import Config, train_mono
    from jperceiver_b200.datasets.get_dataset import get_dataset
    from jperceiver_b200.model import MONO
    path = tmp_path / "cfg_small.py"
    path.write_text(CFG)
    cfg = Config.fromfile(str(path))
    cfg.work_dir = str(tmp_path / "work")
    torch.manual_seed(1024)
    ds_train, ds_val = get_dataset(cfg.data), get_dataset(cfg.data, training=False)
    model = MONO.module_dict[cfg.model["name"]](cfg.model)
    engine = train_mono(model, ds_train, ds_val, cfg, distributed=False, validate=True)
    work = cfg.work_dir
    for name in ("epoch_1.pth", "epoch_2.pth", "latest.pth"):
        assert os.path.exists(os.path.join(work, name)), name
    ck = torch.load(os.path.join(work, "epoch_2.pth"), weights_only=False)
    assert set(ck) == {"meta", "state_dict", "optimizer"} and ck["meta"]["epoch"] == 2 and ck["meta"]["iter"] == 8
    assert len(ck["state_dict"]) == 766 and all(not v.is_cuda for v in ck["state_dict"].values())
    assert ck["optimizer"]["param_groups"][0]["initial_lr"] == 1e-4 and abs(ck["optimizer"]["param_groups"][0]["lr"] - 5e-5) < 1e-12
    logs = [f for f in os.listdir(work) if f.endswith(".log.json")]
    lines = [json.loads(l) for l in open(os.path.join(work, logs[0]))]
    train = [l for l in lines if l["mode"] == "train"]
    val = [l for l in lines if l["mode"] == "val"]
    assert len(train) == 8 and train[0]["lr"] == 1e-4 and abs(train[-1]["lr"] - 5e-5) < 1e-12
    assert all(l["loss"] == l["loss"] for l in train)                    # finite every step
    assert len(val) == 2 and all(k in val[0] for k in ("abs_rel", "a1", "scale mean", "iou_road", "mAP_road"))
    assert 0.0 <= val[0]["a1"] <= 1.0 and val[0]["abs_rel"] > 0.0
    params_after = engine.flat.param.clone()
    # resume in a fresh model/runner and continue one more epoch; the weights it starts from are the checkpoint's
    cfg2 = Config.fromfile(str(path))
    cfg2.work_dir = str(tmp_path / "work_resumed")
    cfg2.resume_from = os.path.join(work, "latest.pth")
    cfg2.total_epochs = 3
    model2 = MONO.module_dict[cfg2.model["name"]](cfg2.model)
    from jperceiver_b200.apis.runner import Runner
    probe = Runner(model2.to("cuda:0"), cfg2.optimizer, cfg2.optimizer_config, cfg2.work_dir)
    probe.resume(cfg2.resume_from)
    assert probe.epoch == 2 and probe.iter == 8
    assert torch.equal(probe.engine.flat.param, params_after)
    assert int(probe.engine.optimizer.step_count.item()) == 8
    engine2 = train_mono(model2, ds_train, ds_val, cfg2, distributed=False, validate=True)
    ck3 = torch.load(os.path.join(cfg2.work_dir, "epoch_3.pth"), weights_only=False)
    assert ck3["meta"]["epoch"] == 3 and ck3["meta"]["iter"] == 12
    assert bool(torch.isfinite(engine2.flat.param).all().item())
    assert not torch.equal(engine2.flat.param, params_after)


def test_graph_replay_follows_a_changed_calibration():
    """The static scale label's quad mask is rasterised on the host from sample 0's calibration and its device buffer is baked into
    the captured step graph: replaying a batch with a DIFFERENT calibration must refresh that buffer (TrainEngine.
    _refresh_host_caches) — the scale loss of the replay equals the eager step's on the same batch."""
    import bench
    from jperceiver_b200 import synthetic
    from jperceiver_b200.apis import TrainEngine, change_input_variable
    from jperceiver_b200.model import MONO
    dev = torch.device("cuda:0")
    opt = bench.model_options(bench.CONFIGS["C2"], 1)          # 320x1024: the two calibrations below give different, non-empty masks
    losses = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(3)
        model = MONO.module_dict["Baseline"](opt).to(dev).train()
        model.DepthDecoder.drop_p = 0.0                        # same disparities in both runs whatever the step counter
        model.noise_scale = 0.0
        engine = TrainEngine(model, dict(type="Adam", lr=0.0, weight_decay=0), dict(max_norm=35, norm_type=2))
        a = synthetic.make_batch(opt, 1, seed=5, pin=True)
        b = synthetic.make_batch(opt, 1, seed=5, pin=True)
        for k in (("odometry_K", 0, 0), ("Tr_cam2_velo", 0, 0)):
            b[k] = b[k].clone()
        b[("odometry_K", 0, 0)][..., 0, 0] *= 1.15            # another sequence: different focal length / lever arm
        b[("Tr_cam2_velo", 0, 0)][..., 0, 3] += 0.35
        da, db = change_input_variable(a, dev), change_input_variable(b, dev)
        if mode == "eager":
            engine.step(da, need_log=True)
            out = engine.step(db, need_log=True)
        else:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                engine.capture(da, warmup=3)
                engine.replay(da)
                out = engine.replay(db).clone()
            side.synchronize()
            engine.release_graph()
        names = engine.last_names
        losses[mode] = {n: float(v) for n, v in zip(names, out.detach().cpu())}
    from jperceiver_b200.model.mono_baseline.net import static_quad_mask_host
    K0, T0 = a[("odometry_K", 0, 0)][0].numpy(), a[("Tr_cam2_velo", 0, 0)][0].numpy()
    K1, T1 = b[("odometry_K", 0, 0)][0].numpy(), b[("Tr_cam2_velo", 0, 0)][0].numpy()
    m0 = static_quad_mask_host(K0, T0, opt["split"], opt["occ_map_size"], 320, 1024)
    m1 = static_quad_mask_host(K1, T1, opt["split"], opt["occ_map_size"], 320, 1024)
    assert m0.sum() > 0 and (m0 != m1).sum() > 0               # the case is not vacuous
    keys = [n for n in losses["eager"] if "scale" in str(n)]
    assert keys
    for n in keys:
        e, g = losses["eager"][n], losses["graph"][n]
        assert abs(e - g) <= 1e-4 * max(abs(e), 1e-6), (n, e, g)
