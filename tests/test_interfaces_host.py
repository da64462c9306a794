"""CPU tests of the drop-in boundary (trainer.get_trainer + fused interfaces) with the torch test
double as kernel backend: FOMAML / multi-task meta-steps against the goldens of the live
reference, Reptile against its definition, and the world-size-2 partition + all-reduce (gloo)."""
import argparse
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from metaasr_crossaccent_b200 import interfaces as I
from metaasr_crossaccent_b200.trainer import get_trainer
from oracle import port
from tests.helpers import GOLD, check_adam_weights, check_summary, load_batch, load_weights, tiny_cfg
from tests.torch_backend import TorchBackend

ID2ACCENT = {"ca": "canada", "en": "england", "hk": "hongkong"}


def make_config(meta=True, k=0.02, warmup=4):
    am = {"idim": 83, "nheads": 4, "d_model": 32, "d_inner": 64, "dropout": 0.0, "tgt_share_weight": 1,
          "encoder": {"nlayers": 2}, "decoder": {"nlayers": 2}, "pos_dropout": 0.0}
    if meta:
        am.update({"inner_optimizer_cls": "SGD", "inner_optimizer_opt": {"momentum": 0.9, "nesterov": True},
                   "meta_opt_cls": "noam", "meta": {"optimizer_opt": {"k": k, "warmup_steps": warmup}}})
    else:
        am.update({"optimizer_cls": "noam", "optimizer_opt": {"k": k, "warmup_steps": warmup}})
    solver = {"setting": "t", "total_steps": 10, "label_smoothing": 0.2, "eval_ival": 100000, "log_ival": 100000,
              "save_ival": 100000, "spm_mapping": "/nonexistent"}
    return {"asr_model": am, "solver": solver}


def make_paras(algo, meta_k=2):
    return argparse.Namespace(pretrain_accents=["ca", "en"], num_pretrain=2, tgt_accent="hk", runs=0, seed=531,
                              meta_k=meta_k, meta_batch_size=2, sample_strategy="normal", max_step=0, resume=False,
                              algo=algo, pretrain_suffix="t", log_root=None,
                              backend_factory=lambda dt: TorchBackend("cpu", dt))


def make_solver(algo, meta=True):
    cls = I.MultiASRInterface if algo == "multi" else I.FOMetaASRInterface
    s = get_trainer(cls, make_config(meta), make_paras(algo), ID2ACCENT)
    s.set_model()
    s.asr_model.load_state_dict(load_weights(tiny_cfg()))
    if algo != "multi":
        s._original_flat.copy_(s.asr_model.engine.params)
    return s


def test_run_batch_contract():
    z = np.load(GOLD / "run_batch_tiny.npz")
    s = make_solver("fomaml")
    x, ilens, ys, olens = load_batch(z, "in.")
    info = s.run_batch(0, x, ilens, ys, olens, train=True)
    assert set(info) == {"loss", "acc"} and isinstance(info["loss"], float)
    assert abs(info["loss"] - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    assert np.array_equal(olens.numpy(), z["olens_after"])
    params = dict(s.asr_model.named_parameters())
    assert len(list(s.asr_model.parameters())) == len(port.trainable_names(tiny_cfg()))
    for n, p in params.items():                      # .grad populated on asr_model.parameters()
        check_summary(z, "g.", n, p.grad, rtol_l2=2e-3, atol_sample=2e-3)
    # reference-style code keeps working on the views
    gn = torch.nn.utils.clip_grad_norm_(s.asr_model.parameters(), 5)
    assert float(gn) > 0
    sd = s.asr_model.state_dict()
    assert list(sd.keys()) == list(port.param_shapes(tiny_cfg()).keys())
    logit, gold = s.asr_model(x, ilens, ys, olens.clone())
    assert np.abs(logit.numpy() - z["logit"]).max() < 2e-5 and np.array_equal(gold.numpy(), z["gold"])


def _sync_port_to_solver(ml, s):
    """Teacher forcing: give the oracle port the solver's meta weights and Adam state."""
    eng = s.asr_model.engine
    st = s.meta_opt.state
    for n in ml.meta_names:
        src = "char_trans.weight" if n == "pre_embed.weight" else n
        ml.original[n].copy_(eng.layout.view(s._original_flat, src))
        ml.meta_opt.m[n] = eng.layout.view(st.m, src).clone()
        ml.meta_opt.v[n] = eng.layout.view(st.v, src).clone()
    ml.meta_opt.t, ml.meta_step_num = st.t, s.meta_opt.step_num


def test_fomaml_meta_steps_match_reference():
    """Step 0 is compared with the live reference's golden; later steps start from weights in which
    Adam(eps=1e-9) has amplified fp32 re-ordering noise of ~zero gradients into +-lr moves,
    so they are compared loosely with the golden and STRICTLY with the oracle port started from the
    solver's own meta weights / Adam state (teacher forcing)."""
    z = np.load(GOLD / "fomaml_tiny.npz")
    s = make_solver("fomaml")
    assert abs(s.inner_lr - float(z["inner_lr"])) < 1e-15
    eng = s.asr_model.engine
    ml = port.MetaLearner(load_weights(tiny_cfg()), tiny_cfg(), algo="fomaml", k=float(z["k"]), warmup=int(z["warmup_steps"]))
    for step in range(int(z["n_meta_steps"])):
        tasks, otasks = [], []
        for acc in range(int(z["n_accents"])):
            tr = [(acc, load_batch(z, f"s{step}.a{acc}.tr{j}.")) for j in range(int(z["meta_k"]))]
            tasks.append((tr, (acc, load_batch(z, f"s{step}.a{acc}.te."))))
            otasks.append(([load_batch(z, f"s{step}.a{acc}.tr{j}.") for j in range(int(z["meta_k"]))],
                           load_batch(z, f"s{step}.a{acc}.te.")))
        _sync_port_to_solver(ml, s)
        captured = {}
        orig_step = s.meta_opt.step

        def spy(upd, count, _o=orig_step, _c=captured):
            _c["mg"] = (upd / count).clone()
            return _o(upd, count)
        s.meta_opt.step = spy
        s.meta_step_on_tasks(tasks)
        s.meta_opt.step = orig_step
        infos = s.flush_train_info()
        oinfos, olr = ml.meta_step(otasks)
        assert abs(s.meta_opt.lr - float(z[f"s{step}.lr"])) < 1e-12 and abs(olr - s.meta_opt.lr) < 1e-15
        strict = step == 0
        for acc, info in enumerate(infos):
            ref = float(z[f"s{step}.a{acc}.te_loss"])
            assert abs(info["loss"] - ref) <= (2e-4 if strict else 2e-2) * abs(ref)
            assert abs(info["loss"] - oinfos[acc]["loss"]) <= 2e-4 * abs(ref)
        for n in eng.layout.shapes:
            if n == "pos_encoder.pe":
                continue
            src = "char_trans.weight" if n == "pre_embed.weight" else n
            mg = eng.layout.view(captured["mg"], src)
            check_summary(z, f"s{step}.mg.", n, mg, rtol_l2=5e-3 if strict else 5e-2, atol_sample=3e-2 if strict else 0.3)
            omg = ml.last_meta_grad[n]
            assert float((mg - omg).norm()) <= 5e-3 * float(omg.norm()) + 1e-9, (step, n)
            if strict:
                check_adam_weights(z, f"s{step}.w.", ["s0.mg."], n, s._original[n], s.meta_opt.lr)
            solid = omg.abs() > max(1e-7, 0.05 * float(omg.abs().max()))
            err = (s._original[n] - ml.original[n]).abs()[solid]
            assert err.numel() == 0 or float(err.max()) <= 3e-2 * s.meta_opt.lr, (step, n)


def test_multi_steps_match_reference():
    z = np.load(GOLD / "multi_tiny.npz")
    s = make_solver("multi", meta=False)
    for step in range(int(z["n_steps"])):
        info = s.multi_step((0, load_batch(z, f"s{step}.")))
        ref = float(z[f"s{step}.loss"])
        assert abs(info["loss"] - ref) <= 5e-4 * abs(ref)
        lr = port.noam_lr(step + 1, float(z["k"]), 32, int(z["warmup_steps"]))
        assert abs(s.asr_opt.lr - lr) < 1e-15
        for n, t in s.asr_model.state_dict().items():
            if n not in ("pos_encoder.pe", "pre_embed.weight"):
                check_adam_weights(z, f"s{step}.w.", [f"s{i}.g." for i in range(step + 1)], n, t, lr)


def test_reptile_matches_definition():
    z = np.load(GOLD / "fomaml_tiny.npz")
    s = make_solver("reptile")
    ml = port.MetaLearner(load_weights(tiny_cfg()), tiny_cfg(), algo="reptile", k=0.02, warmup=4)
    tasks, otasks = [], []
    for acc in range(2):
        tr = [load_batch(z, f"s0.a{acc}.tr{j}.") for j in range(2)]
        te = load_batch(z, f"s0.a{acc}.te.")
        otasks.append(([tuple(t if not isinstance(t, list) else [y.clone() for y in t] for t in b) for b in tr], te))
        tasks.append(([(acc, load_batch(z, f"s0.a{acc}.tr{j}.")) for j in range(2)], (acc, load_batch(z, f"s0.a{acc}.te."))))
    s.meta_step_on_tasks(tasks)
    ml.meta_step(otasks)
    for n in ml.meta_names:
        g = ml.last_meta_grad[n]
        mask = g.abs() > 1e-7
        err = (s._original[n] - ml.original[n]).abs()[mask]
        assert err.numel() == 0 or float(err.max()) <= 3e-2 * s.meta_opt.lr, n


# ---------------------------------------------------------------------------- world size 2 (gloo)
def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port_no, out):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port_no)})
    torch.set_num_threads(2)
    from metaasr_crossaccent_b200 import dist as D
    D.init_from_env("gloo")
    z = np.load(GOLD / "fomaml_tiny.npz")
    s = make_solver("fomaml")
    mine = D.partition_tasks([0, 1], 2)
    tasks = []
    for acc in mine:
        tr = [(acc, load_batch(z, f"s0.a{acc}.tr{j}.")) for j in range(2)]
        tasks.append((tr, (acc, load_batch(z, f"s0.a{acc}.te."))))
    s.meta_step_on_tasks(tasks, global_task_count=2)
    torch.save(s._original_flat.clone(), f"{out}/w{rank}.pt")
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_meta_step_equals_sequential(tmp_path):
    """Accents partitioned over 2 ranks + one all-reduce == the sequential reference loop."""
    port_no = _free_port()
    mp.spawn(_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    w0, w1 = torch.load(tmp_path / "w0.pt"), torch.load(tmp_path / "w1.pt")
    assert torch.equal(w0, w1)                                   # replicas stay in sync
    z = np.load(GOLD / "fomaml_tiny.npz")
    s = make_solver("fomaml")
    tasks = []
    for acc in range(2):
        tr = [(acc, load_batch(z, f"s0.a{acc}.tr{j}.")) for j in range(2)]
        tasks.append((tr, (acc, load_batch(z, f"s0.a{acc}.te."))))
    s.meta_step_on_tasks(tasks)
    lr = s.meta_opt.lr
    eng = s.asr_model.engine
    for n in eng.layout.offsets:
        check_adam_weights(z, "s0.w.", ["s0.mg."], n, eng.layout.view(w0, n), lr)
        ref = eng.layout.view(s._original_flat, n)
        mask = torch.from_numpy(np.abs(z[f"s0.mg.{n}#sample"]) > 1e-7)
    assert float((w0 - s._original_flat).abs().max()) <= 2.0 * lr + 1e-12


def _multi_worker(rank, world, port_no, out):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port_no)})
    torch.set_num_threads(2)
    from metaasr_crossaccent_b200 import dist as D
    D.init_from_env("gloo")
    z = np.load(GOLD / "multi_tiny.npz")
    s = make_solver("multi", meta=False)
    infos = []
    for step in range(2):                      # rank r trains on batch s{2*step + r} of the golden file's inputs
        infos.append(s.multi_step((0, load_batch(z, f"s{(2 * step + rank) % 3}."))))
    torch.save({"w": s.asr_model.engine.params.clone(), "loss": [i["loss"] for i in infos],
                "lr": s.asr_opt.lr, "step_num": s.asr_opt.step_num}, f"{out}/m{rank}.pt")
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_multi_task_gradient_dp_equals_mean_gradient_step(tmp_path):
    """Multi-task data parallelism (loop body multi_interface.py:100-114): two ranks draw
    different batches, one all-reduce of the flat gradient arena, then clip_grad_norm_(5) and noam-Adam on the
    MEAN gradient.  Replicas end bit-identical and equal a single process stepping on the averaged gradient."""
    port_no = _free_port()
    mp.spawn(_multi_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "m0.pt"), torch.load(tmp_path / "m1.pt")
    assert torch.equal(r0["w"], r1["w"]) and r0["lr"] == r1["lr"] and r0["step_num"] == r1["step_num"] == 2
    z = np.load(GOLD / "multi_tiny.npz")
    s = make_solver("multi", meta=False)
    eng = s.asr_model.engine
    n = eng.layout.total
    for step in range(2):
        losses, gsum = [], torch.zeros_like(eng.grads)
        for r in range(2):
            info = s.run_batch(0, *load_batch(z, f"s{(2 * step + r) % 3}."), train=True, accent_idx=0)
            losses.append(info["loss"])
            gsum += eng.grads
        eng.grads.copy_(gsum / 2)
        eng.be.mt_sumsq(eng.grads[:n], s._gnorm)
        s.asr_opt.step(s._gnorm, I.GRAD_CLIP)
        assert abs(losses[0] - r0["loss"][step]) <= 1e-6 * abs(losses[0])      # each rank saw its own batch
        assert abs(losses[1] - r1["loss"][step]) <= 1e-6 * abs(losses[1])
    assert abs(s.asr_opt.lr - r0["lr"]) < 1e-15
    # Adam(eps 1e-9) amplifies fp32 re-ordering noise on ~0 gradients to +-lr: compare in units of lr
    diff = (eng.params - r0["w"]).abs()
    assert float(diff.max()) <= 2.0 * s.asr_opt.lr + 1e-12
    assert float((diff > 3e-2 * s.asr_opt.lr).float().mean()) < 0.02


def test_recog_greedy_ids_bit_exact_host_logic():
    """MyTransformer.recog (greedy decode on the growing prefix, encoder memory computed once) through the engine's
    schedule on the torch test double: token ids bit-exact against the live reference's golden."""
    z = np.load(GOLD / "run_batch_tiny.npz")
    s = make_solver("fomaml")
    x, ilens, _, _ = load_batch(z, "in.")
    ids = s.asr_model.recog(x, ilens)                       # key/value-cached: one decoder row per step
    assert np.array_equal(ids.numpy(), z["greedy"])
    ids_ref_schedule = s.asr_model.recog(x, ilens, kv_cache=False)      # the reference's O(L^2) re-run schedule
    assert np.array_equal(ids_ref_schedule.numpy(), z["greedy"])


# ---------------------------------------------------------------------------- fine-tune loop
def _mono_solver(pre_path, tmp_path, backend_factory, **am_extra):
    from tests.helpers import mono_paras
    cfg = make_config(meta=False)
    cfg["asr_model"].update(am_extra)
    cfg["solver"].update({"pretrain_module": ["feat_extractor", "vgg2enc", "encoder"], "freeze_module": ["encoder"],
                          "total_epochs": 2})
    paras = mono_paras(tmp_path, pre_path, backend_factory=backend_factory)
    s = get_trainer(I.MonoASRInterface, cfg, paras, ID2ACCENT)
    return s


def test_mono_finetune_filter_freeze_steps_match_reference(tmp_path):
    """MonoASRInterface (mono_interface.py:75-148): filter_model over pretrain_module, freeze_module, then
    run_batch -> clip -> noam-Adam, against the golden made with the reference's own methods."""
    from tests.helpers import run_mono_freeze_check, set_model_from_tiny_init
    z = np.load(GOLD / "mono_freeze_tiny.npz")

    def mk(pre_path):
        return set_model_from_tiny_init(_mono_solver(pre_path, tmp_path, lambda dt: TorchBackend("cpu", dt)))
    s = run_mono_freeze_check(mk, z, tmp_path, loss_rtol=5e-4)
    # per-epoch checkpoint + resume: files of mono_interface.py:34-59, state restored exactly
    s.train_set = [load_batch(z, "s0."), load_batch(z, "s1.")]
    s.max_epoch = 1
    s.train()
    for f in ("snapshot.latest", "optimizer.latest", "info_dict.latest", "global_step", "epoch"):
        assert s.log_dir.joinpath(f).exists(), f
    sd = torch.load(s.log_dir.joinpath("snapshot.latest"))
    assert list(sd.keys()) == list(s.asr_model.state_dict().keys())
    from tests.helpers import mono_paras
    cfg = make_config(meta=False)
    cfg["solver"].update({"pretrain_module": ["encoder"], "total_epochs": 2})
    paras = mono_paras(tmp_path, tmp_path / "snapshot.step.100", resume=True,
                       backend_factory=lambda dt: TorchBackend("cpu", dt))
    r = get_trainer(I.MonoASRInterface, cfg, paras, ID2ACCENT)
    r.set_model()
    assert r.ep == 1 and r.global_step == s.global_step and r.asr_opt.step_num == s.asr_opt.step_num
    for (n, a), (_, b) in zip(r.asr_model.state_dict().items(), s.asr_model.state_dict().items()):
        assert torch.equal(a, b), n
    assert torch.equal(r.asr_opt.state.m, s.asr_opt.state.m) and torch.equal(r.asr_opt.state.v, s.asr_opt.state.v)


def test_next_tasks_handover_is_a_no_op_on_the_host_double():
    """meta_step_on_tasks(next_tasks=) (copy-stream prefetch on the CUDA path) must not change the schedule: on the CPU double
    staging passes batches through, the handed-over list is picked up by identity at the next call, and two steps with the
    hand-over equal two plain steps bit for bit; optimizer_state() stays a local (non-collective) call."""
    z = np.load(GOLD / "fomaml_tiny.npz")

    def tasks_of():
        out = []
        for acc in range(int(z["n_accents"])):
            tr = [(acc, load_batch(z, f"s0.a{acc}.tr{j}.")) for j in range(int(z["meta_k"]))]
            out.append((tr, (acc, load_batch(z, f"s0.a{acc}.te."))))
        return out
    res = []
    for handover in (False, True):
        s = make_solver("fomaml")
        t1, t2 = tasks_of(), tasks_of()
        s.meta_step_on_tasks(t1, next_tasks=t2 if handover else None)
        assert (s.__dict__.get("_prefetched") is not None) == handover
        s.meta_step_on_tasks(t2)
        assert s.__dict__.get("_prefetched") is None and getattr(s, "_nvls", None) is None
        osd = s.optimizer_state()
        res.append((s._original_flat.clone(), osd["m"].clone(), [i["loss"] for i in s.flush_train_info()]))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
