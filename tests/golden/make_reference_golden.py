"""Generates tests/golden/reference_*.npz / .json by running the REFERENCE'S OWN CODE, unmodified, from where it lies
under /root/reference (read-only), in this container.  Run:  python tests/golden/make_reference_golden.py

The reference's modules import third-party packages that are not installed here (pytorch_lightning, lhotse,
torchmetrics, ml_collections, asteroid_filterbanks, wandb ...).  None of them contributes arithmetic to
the functions pinned below -- they provide a base class, metric objects, a config container and manifest types -- so
they are replaced by inert stub modules installed in sys.modules before the import:
  * pytorch_lightning.LightningModule -> nn.Module + `save_hyperparameters(*names)` (captures the named locals of the
    calling __init__ into `self.hparams`, as Lightning does) + no-op `log` / `log_dict`;
  * ml_collections.ConfigDict -> a dict with attribute access;
  * everything else (lhotse.*, torchmetrics.*, asteroid_filterbanks.*, wandb, s3prl ...) -> modules whose every attribute
    is a dummy class.
Two more shims, both outside the arithmetic:
  * `median_filter` ends with a hard-coded `.to("cuda")` (src/utils/helper.py:95); there is no GPU here, so while the
    reference runs `torch.Tensor.to` maps the device "cuda" to "cpu";
  * the run-length extraction of `get_new_cuts` (src/scripts/predict.py:472-490) is a loop inside a 200-line function
    that needs lhotse cut objects; its `for k, value in enumerate(obj["tensor"])` statement and the trailing-run block
    are lifted from the file's syntax tree and executed verbatim on a seeded stream.  `merge_intervals_with_buffer`,
    `split_into_windows`, `get_binary_tensor`, `get_false_alarm`, `get_missed_detection` (predict.py:614-673) and
    `get_timestamp_from_sample_boundary` (predict_sincnet.py:492-504) are function definitions compiled from the
    reference files' syntax trees the same way (importing those scripts would pull in the data modules).
What is pinned: PyanNet2 construction order / forward (a2), VadModel.forward / predict_step (a6), median_filter (a7),
the RLE and its seconds arithmetic (a8, a8'), merge / split (a10), the frame arithmetic (a5), load_config (a12) and the
DER helpers (f1).  What is NOT (arithmetic lives in absent third-party code): lhotse Fbank (a1) and asteroid's
ParamSincFB filters inside SincNet (a3 / a4).

The fixtures cannot be regenerated on the GPU box (/root/reference does not exist there); tests only read them.
"""
import ast
import hashlib
import importlib.abc
import importlib.machinery
import inspect
import json
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
STUB_ROOTS = ("pytorch_lightning", "lhotse", "torchmetrics", "ml_collections", "asteroid_filterbanks", "wandb", "s3prl",
              "lightning", "jiwer", "whisper", "pyannote", "speechbrain", "torchaudio_stub")


class _HParams(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class LightningModule(nn.Module):
    """nn.Module + the two Lightning services the reference models use."""

    def save_hyperparameters(self, *names):
        frame = inspect.currentframe().f_back
        if not hasattr(self, "_hp"):
            object.__setattr__(self, "_hp", _HParams())
        for n in names:
            self._hp[n] = frame.f_locals[n]

    @property
    def hparams(self):
        return self._hp

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = dict.__setitem__


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if self.__name__ == "pytorch_lightning" and name == "LightningModule":
            return LightningModule
        if self.__name__ == "ml_collections" and name == "ConfigDict":
            return ConfigDict
        cls = type(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _lift(path, names):
    """Compile the named top-level function definitions of a reference file, unmodified, into a namespace."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "math": math, "np": np}
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert {n.name for n in body} == set(names), (path, names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def _lift_rle(path):
    """The RLE statements of get_new_cuts (predict.py:471-490): `start = None`, the frame loop, the trailing-run block."""
    tree = ast.parse(open(path).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "get_new_cuts")
    for node in ast.walk(fn):
        if not isinstance(node, ast.For):
            continue
        for i, st in enumerate(node.body):
            if (isinstance(st, ast.For) and isinstance(st.iter, ast.Call) and getattr(st.iter.func, "id", "") == "enumerate"
                    and "obj" in ast.unparse(st.iter) and "tensor" in ast.unparse(st.iter)):
                stmts = [node.body[i - 1], st, node.body[i + 1]]
                assert ast.unparse(stmts[0]).strip() == "start = None" and isinstance(stmts[2], ast.If)
                return compile(ast.Module(body=stmts, type_ignores=[]), path, "exec"), (stmts[0].lineno, stmts[2].end_lineno)
    raise RuntimeError("RLE loop not found")


def state_hash(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def spread(vm, x, sigma=1.0):
    """Rescale the classifier so that the logits on x have zero mean and standard deviation sigma (1: the original fixtures;
    2 and 4: what a trained detector looks like, p from ~1e-4 to ~1 - 1e-4)."""
    p = vm(x).double()
    z = torch.log(p / (1 - p))
    scale = (float(sigma) / z.std().clamp_min(1e-9)).float()
    cls = vm.model.classifier
    cls.bias.copy_((cls.bias - z.mean().float()) * scale)
    cls.weight.mul_(scale)


def main():
    assert os.path.isdir(REF), "the reference tree is needed to regenerate these fixtures"
    torch.set_num_threads(1)
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REF)
    orig_to = torch.Tensor.to

    def to_cpu_for_cuda(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return orig_to(self, *a, **k)

    # asteroid-filterbanks (requirements.txt:1) is absent: the reference's SincNet gets the oracle's restatement of
    # Encoder / ParamSincFB as its filterbank, so that everything ELSE in sincnet.py / PyanNet.py (the convolution stack, |x| on the
    # sinc layer only, MaxPool3 -> InstanceNorm -> LeakyReLU order, the "b f t -> b t f" rearrange, the head) is the reference's own
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import oracle.models as oracle_models
    afb = types.ModuleType("asteroid_filterbanks")
    afb.Encoder, afb.ParamSincFB = oracle_models.Encoder, oracle_models.ParamSincFB
    sys.modules["asteroid_filterbanks"] = afb

    from src.engines.vad_engine import VadModel                      # the reference's own modules
    from src.models.segmentation.PyanNet2 import PyanNet2
    from src.utils.helper import median_filter
    from src.utils import receptive_field as rf
    from config.config import load_config

    out, meta = {}, {"reference_files": {}}

    # ---- a2 / a6: PyanNet2 through VadModel, seed 42 (config/config.py:11), fbank-dim input
    g = torch.Generator().manual_seed(123)
    feats = torch.randn(3, 120, 80, generator=g) * 3 - 5
    labels = (torch.rand(3, 120, generator=g) > 0.5).float()
    torch.manual_seed(42)
    vm = VadModel("PyanNet2", {"encoding_dim": 80}).eval()
    meta["pyannet2_d80_seed42_state_sha256"] = state_hash(vm.state_dict())
    meta["pyannet2_d80_state_keys"] = {k: list(v.shape) for k, v in vm.state_dict().items()}
    with torch.no_grad():
        out["d80_feats"] = feats.numpy()
        out["d80_labels"] = labels.numpy()
        out["d80_prob"] = vm(feats).numpy()
        torch.Tensor.to = to_cpu_for_cuda
        try:
            out["d80_predict"] = vm.predict_step({"inputs": feats, "is_voice": labels}, 0).numpy()
        finally:
            torch.Tensor.to = orig_to
    # the same with a classifier rescaled (a parameter change, not a code change) so that the logits are ~N(0, 1): the
    # decisions of a random-init model are otherwise constant
    with torch.no_grad():
        spread(vm, feats)
        out["d80_spread_cls_w"] = vm.model.classifier.weight.numpy().copy()
        out["d80_spread_cls_b"] = vm.model.classifier.bias.numpy().copy()
        out["d80_spread_prob"] = vm(feats).numpy()
        torch.Tensor.to = to_cpu_for_cuda
        try:
            out["d80_spread_predict"] = vm.predict_step({"inputs": feats, "is_voice": labels}, 0).numpy()
        finally:
            torch.Tensor.to = orig_to

    # trained-model logit spreads: the same model and inputs with the classifier rescaled to sigma = 2 and 4
    for sg in (2, 4):
        with torch.no_grad():
            spread(vm, feats, sg)
            out[f"d80_s{sg}_cls_w"] = vm.model.classifier.weight.numpy().copy()
            out[f"d80_s{sg}_cls_b"] = vm.model.classifier.bias.numpy().copy()
            out[f"d80_s{sg}_prob"] = vm(feats).numpy()
            torch.Tensor.to = to_cpu_for_cuda
            try:
                out[f"d80_s{sg}_predict"] = vm.predict_step({"inputs": feats, "is_voice": labels}, 0).numpy()
            finally:
                torch.Tensor.to = orig_to
    # BASELINE config 2's row shape (8 s = 800 frames x 80 mel bins), one row, sigma = 2 classifier: the reference's own output
    g2 = torch.Generator().manual_seed(321)
    feats8 = torch.randn(1, 800, 80, generator=g2) * 3 - 5
    with torch.no_grad():
        spread(vm, feats8, 2)
        out["cfg2_cls_w"] = vm.model.classifier.weight.numpy().copy()
        out["cfg2_cls_b"] = vm.model.classifier.bias.numpy().copy()
        out["cfg2_feats"] = feats8.numpy()
        out["cfg2_prob"] = vm(feats8).numpy()
        torch.Tensor.to = to_cpu_for_cuda
        try:
            out["cfg2_predict"] = vm.predict_step({"inputs": feats8, "is_voice": torch.zeros(1, 800)}, 0).numpy()
        finally:
            torch.Tensor.to = orig_to

    # ---- a2: a small PyanNet2 with every weight stored (independent of the RNG stream): monolithic and layer-wise LSTMs
    for tag, mono in (("tiny_mono", True), ("tiny_split", False)):
        torch.manual_seed(7)
        m = PyanNet2(lstm={"hidden_size": 8, "num_layers": 2, "monolithic": mono}, linear={"hidden_size": 6, "num_layers": 2},
                     encoding_dim=5)
        m.build()
        m.eval()
        x = torch.randn(2, 17, 5, generator=g)
        with torch.no_grad():
            out[f"{tag}_x"] = x.numpy()
            out[f"{tag}_prob"] = m(x).numpy()
        for k, v in m.state_dict().items():
            out[f"{tag}_w_{k}"] = v.numpy()

    # ---- a3 / a4: the reference's PyanNet / SincNet classes around the restated filterbank
    torch.manual_seed(42)
    vmp = VadModel("PyanNet", {}).eval()
    meta["pyannet_seed42_state_sha256"] = state_hash(vmp.state_dict())
    meta["pyannet_state_keys"] = {k: list(v.shape) for k, v in vmp.state_dict().items()}
    wavp = 0.1 * torch.randn(2, 16000, generator=g)
    wavp[1] *= torch.linspace(0.0, 1.0, 16000)
    with torch.no_grad():
        out["pyannet_wav"] = wavp.numpy()
        out["pyannet_sincnet"] = vmp.model.sincnet(wavp.unsqueeze(1)).numpy()
        out["pyannet_prob"] = vmp.model(wavp.unsqueeze(1)).numpy()
        torch.Tensor.to = to_cpu_for_cuda
        try:
            out["pyannet_predict"] = vmp.predict_step({"inputs": wavp, "is_voice": torch.zeros(2, out["pyannet_prob"].shape[1])}, 0).numpy()
        finally:
            torch.Tensor.to = orig_to

    # ---- a6: SSL-dim model -> median window 25 (vad_engine.py:207)
    torch.manual_seed(42)
    vm768 = VadModel("PyanNet2", {"encoding_dim": 768}).eval()
    meta["pyannet2_d768_seed42_state_sha256"] = state_hash(vm768.state_dict())
    x768 = torch.randn(2, 60, 768, generator=g)
    with torch.no_grad():
        spread(vm768, x768)
        out["d768_cls_w"] = vm768.model.classifier.weight.numpy().copy()
        out["d768_cls_b"] = vm768.model.classifier.bias.numpy().copy()
        out["d768_x"] = x768.numpy()
        out["d768_prob"] = vm768(x768).numpy()
        torch.Tensor.to = to_cpu_for_cuda
        try:
            out["d768_predict"] = vm768.predict_step({"inputs": x768, "is_voice": torch.zeros(2, 60)}, 0).numpy()
        finally:
            torch.Tensor.to = orig_to
    for sg in (2, 4):
        with torch.no_grad():
            spread(vm768, x768, sg)
            out[f"d768_s{sg}_cls_w"] = vm768.model.classifier.weight.numpy().copy()
            out[f"d768_s{sg}_cls_b"] = vm768.model.classifier.bias.numpy().copy()
            out[f"d768_s{sg}_prob"] = vm768(x768).numpy()

    # ---- a7: median_filter on its own (helper.py:66-97), both windows, ties / NaN / short rows
    prob = torch.sigmoid(torch.cumsum(torch.randn(5, 400, generator=g), 1) * 0.3)
    prob[0, 100] = 0.5
    prob[1, 7] = float("nan")
    short = torch.rand(3, 10, generator=g)
    torch.Tensor.to = to_cpu_for_cuda
    try:
        out["mf_prob"] = prob.numpy()
        out["mf_49"] = median_filter(prob.clone(), window=0.01).numpy()
        out["mf_25"] = median_filter(prob.clone(), window=0.02).numpy()
        out["mf_default"] = median_filter(prob.clone()).numpy()
        out["mf_short_prob"] = short.numpy()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out["mf_short_49"] = median_filter(short.clone(), window=0.01).numpy()
    finally:
        torch.Tensor.to = orig_to

    # ---- a5: frame arithmetic (receptive_field.py)
    ns = [0, 250, 251, 260, 991, 1261, 16000, 80000, 128000, 960000, 57600000] + [int(v) for v in torch.randint(251, 200000, (20,), generator=g)]
    out["rf_num_samples"] = np.array(ns, dtype=np.int64)
    out["rf_num_frames"] = np.array([rf.get_num_frames(n) if n >= 251 else -1 for n in ns], dtype=np.int64)
    out["rf_field_size"] = np.array([rf.receptive_field_size(k) for k in (1, 2, 3, 10, 293)], dtype=np.int64)
    out["rf_conv1d"] = np.array([rf.conv1d_num_frames(n, 5, 1) for n in (5, 6, 100)] + [rf.conv1d_num_frames(n, 251, 10) for n in (251, 1000)],
                                dtype=np.int64)

    KS, ST, PD, DL = [251, 3, 5, 3, 5, 3], [10, 3, 1, 3, 1, 3], [0] * 6, [1] * 6
    out["rf_multi_frames"] = np.array([rf.multi_conv_num_frames(n, kernel_size=KS, stride=ST, padding=PD, dilation=DL) for n in (991, 16000, 80000)], dtype=np.int64)
    out["rf_multi_size"] = np.array([rf.multi_conv_receptive_field_size(k, kernel_size=KS, stride=ST, dilation=DL) for k in (1, 2, 471)], dtype=np.int64)
    out["rf_conv_size"] = np.array([rf.conv1d_receptive_field_size(k, kernel_size=5, stride=3, dilation=1) for k in (1, 2, 10)], dtype=np.int64)
    out["rf_conv_center"] = np.array([rf.conv1d_receptive_field_center(f, kernel_size=251, stride=10, padding=0, dilation=1) for f in (0, 1, 100)], dtype=np.int64)
    out["rf_multi_center"] = np.array([rf.multi_conv_receptive_field_center(f, kernel_size=KS, stride=ST, padding=PD, dilation=DL) for f in (0, 1, 292)], dtype=np.int64)

    # ---- a8 / a8' / a10 / f1: predict.py and predict_sincnet.py pieces
    pred_py = os.path.join(REF, "src/scripts/predict.py")
    psinc_py = os.path.join(REF, "src/scripts/predict_sincnet.py")
    fns = _lift(pred_py, ["merge_intervals_with_buffer", "split_into_windows", "get_binary_tensor", "get_false_alarm", "get_missed_detection"])
    ts = _lift(psinc_py, ["get_timestamp_from_sample_boundary"])["get_timestamp_from_sample_boundary"]
    rle_code, rle_lines = _lift_rle(pred_py)
    meta["reference_files"]["rle_lines_predict_py"] = list(rle_lines)
    streams = []
    for i in range(6):
        s = (torch.cumsum(torch.randn(700, generator=g), 0) > 0).long()
        if i == 1:
            s[-5:] = 1          # trailing run
        if i == 2:
            s[:] = 0
        if i == 3:
            s[:] = 1
        if i == 4:
            s = torch.tensor([0, 1, 0, 1, 1, 0, 1, 1, 1, 0, 0, 1], dtype=torch.long)   # runs of 1, 2, 3 and a trailing single
        streams.append(s)
    for i, s in enumerate(streams):
        for fs_tag, fs in (("10ms", 0.01), ("20ms", 0.02)):
            loc = {"obj": {"tensor": s}, "frame_shift": fs, "pred_intervals": [], "round": round, "len": len, "enumerate": enumerate}
            exec(rle_code, loc)
            out[f"rle{i}_{fs_tag}_stream"] = s.numpy().astype(np.uint8)
            out[f"rle{i}_{fs_tag}_intervals"] = np.array(loc["pred_intervals"], dtype=np.float64).reshape(-1, 2)
    iv = [(5.2, 7.9), (0.3, 1.1), (1.05, 2.0), (30.0, 55.5), (7.9, 8.0), (59.5, 60.0)]
    for b in (0, 0.25, 1.0):
        out[f"merge_b{b}"] = np.array(fns["merge_intervals_with_buffer"](list(iv), 60.0, b), dtype=np.float64).reshape(-1, 2)
    out["merge_in"] = np.array(iv, dtype=np.float64)
    out["split10"] = np.array(fns["split_into_windows"]([[0.0, 35.05], [40.0, 40.05], [50.0, 60.1], [70.0, 80.0]], window=10), dtype=np.float64)
    gt = fns["get_binary_tensor"]([(0.3, 1.1), (5.2, 7.9)], 10.0, 0.01)
    pr = fns["get_binary_tensor"]([(0.5, 1.5), (5.0, 7.0), (9.0, 9.5)], 10.0, 0.01)
    out["der_gt"], out["der_pred"] = gt.numpy(), pr.numpy()
    out["der_fa_md"] = np.array([float(fns["get_false_alarm"](gt, pr)), float(fns["get_missed_detection"](gt, pr))], dtype=np.float64)
    out["sinc_ts_in"] = np.array([(0, 10, 5), (3, 293, 5), (100, 3551, 60), (0, 0, 1), (59, 118, 2)], dtype=np.int64)
    out["sinc_ts_out"] = np.array([ts(a, b, d) for a, b, d in out["sinc_ts_in"].tolist()], dtype=np.int64)

    # ---- a8-a10 + f1 end to end: the reference's get_new_cuts (predict.py:412-612) on a synthetic manifest pair.
    # lhotse's load_manifest_lazy is I/O glue (one JSON object per line): a 10-line reader stands in; per-recording values are
    # captured by wrapping (not changing) the reference's helpers.
    import contextlib
    import gzip
    import io
    import src.scripts.predict as P

    class _Obj(dict):
        def __getattr__(self, k):
            v = self[k]
            return [_Obj(x) for x in v] if k == "supervisions" else v

        def to_dict(self):
            return dict(self)

    def _load(path):
        with gzip.open(path, "rt") as f:
            return [_Obj(json.loads(line)) for line in f if line.strip()]

    P.load_manifest_lazy = _load
    mdir = os.path.join(HERE, "manifests")
    os.makedirs(mdir, exist_ok=True)
    durs = [4.99, 13.0, 0.5, 20.2, 7.77, 31.4]
    recs, cuts = [], []
    for i, d in enumerate(durs):
        rec = {"id": f"rec{i}", "sources": [{"type": "file", "channels": [0], "source": f"rec{i}.wav"}], "sampling_rate": 16000,
               "num_samples": int(round(d * 16000)), "duration": d, "channel_ids": [0]}
        sups, t = [], 0.0
        while True:
            t += float(torch.rand(1, generator=g)) * 2.0
            du = 0.2 + float(torch.rand(1, generator=g)) * 3.0
            if t + du > d:
                break
            sups.append({"id": f"rec{i}-sup{len(sups)}", "recording_id": f"rec{i}", "start": round(t, 3), "duration": round(du, 3),
                         "channel": 0, "text": f"utt {len(sups)}"})
            t += du
        recs.append(rec)
        cuts.append({"id": f"rec{i}-0", "start": 0, "duration": d, "channel": 0, "supervisions": sups, "recording": rec, "type": "MonoCut"})
    for name, items in (("recordings.jsonl.gz", recs), ("cuts.jsonl.gz", cuts)):
        with open(os.path.join(mdir, name), "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:   # reproducible bytes
            for it in items:
                f.write((json.dumps(it) + "\n").encode())
    nfr = sum(math.ceil(d / 0.01) + 1 for d in durs)
    rows = (nfr + 499) // 500
    preds = (torch.cumsum(torch.randn(rows * 500, generator=g), 0).reshape(rows, 500, 1) > 0).long()
    out["gnc_preds"] = preds.numpy().astype(np.uint8)
    orig = {k: getattr(P, k) for k in ("get_false_alarm", "get_missed_detection", "merge_intervals_with_buffer", "split_into_windows")}
    meta["get_new_cuts"] = {}
    for tag, buffer, split in (("b0", 0, False), ("b025_split", 0.25, True)):
        log = {"fa": [], "md": [], "merged": [], "split": []}
        P.get_false_alarm = lambda a, b, log=log: (log["fa"].append(float(orig["get_false_alarm"](a, b))), orig["get_false_alarm"](a, b))[1]
        P.get_missed_detection = lambda a, b, log=log: (log["md"].append(float(orig["get_missed_detection"](a, b))), orig["get_missed_detection"](a, b))[1]
        P.merge_intervals_with_buffer = lambda iv, d, b, log=log: (lambda r: (log["merged"].append([list(x) for x in r]), r)[1])(orig["merge_intervals_with_buffer"](iv, d, b))
        P.split_into_windows = lambda iv, window=10, log=log: (lambda r: (log["split"].append([list(x) for x in r]), r)[1])(orig["split_into_windows"](iv, window=window))
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            P.get_new_cuts("synthetic", "test", preds, os.path.join(mdir, "recordings.jsonl.gz"), os.path.join(mdir, "cuts.jsonl.gz"),
                           mdir, "unused.jsonl.gz", buffer=buffer, split=split, frame_shift=0.01)
        for k, v in orig.items():
            setattr(P, k, v)
        meta["get_new_cuts"][tag] = {"buffer": buffer, "split": split, "fa": log["fa"], "md": log["md"],
                                     "intervals": log["split"] if split else log["merged"], "report": buf.getvalue()}
    meta["get_new_cuts"]["durations"] = durs

    # ---- the SincNet variant: predict_sincnet.get_new_cuts (predict_sincnet.py:294-489): slices of ceil(get_num_frames(16000 d)) + 1
    # frames, RLE in frame indices, get_timestamp_from_sample_boundary (whole seconds), scoring at 20 ms.  Its tail builds lhotse cuts
    # (cut.truncate / CutSet.to_file): inert fakes stand in for those objects, which carry no arithmetic.
    import tempfile
    import src.scripts.predict_sincnet as PS

    # lhotse's cut objects: a functional stand-in.  ``truncate`` is the package's restatement of lhotse's MonoCut.truncate
    # (b200vad/manifests.py: parity with lhotse itself unpinned, lhotse is absent); everything the REFERENCE does with the
    # truncated cuts (:391-467: counters, text folding, first-supervision rewrite, ids, empty-cut drop) runs for real and
    # the CutSet it would write is captured.
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "universal-voice-activity-detection_b200"))
    from b200vad import manifests as _mf

    class _Cut:
        def __init__(self, d):
            self._d = {k: v for k, v in d.items() if k != "supervisions"}
            self.id, self.start, self.duration = d["id"], d["start"], d["duration"]
            self.supervisions = [types.SimpleNamespace(**dict({"text": None}, **x)) for x in d["supervisions"]]

        def to_dict(self):
            return dict(self._d, id=self.id, start=self.start, duration=self.duration,
                        supervisions=[{k: v for k, v in vars(x).items()} for x in self.supervisions])

        def index_supervisions(self, **k):
            return None

        def truncate(self, offset=0.0, duration=None, keep_excessive_supervisions=True, _supervisions_index=None):
            return _Cut(_mf.truncate_cut(self.to_dict(), offset=offset, duration=duration,
                                         keep_excessive_supervisions=keep_excessive_supervisions))

        def with_id(self, i):
            self.id = i
            return self

    captured = {}

    class _CutSet:
        def __init__(self, cuts):
            self.cuts = cuts

        @classmethod
        def from_cuts(cls, cuts):
            return cls(list(cuts))

        def to_file(self, path):
            captured["cuts"] = [c.to_dict() for c in self.cuts]

    def _load_s(path):
        with gzip.open(path, "rt") as f:
            objs = [json.loads(line) for line in f if line.strip()]
        return [(_Cut(o) if "supervisions" in o else _Obj(o)) for o in objs]

    PS.load_manifest_lazy = _load_s
    PS.CutSet = _CutSet
    sdurs = [4.99, 13.0, 1.5, 20.2, 7.77, 31.4]
    srecs, scuts = [], []
    for i, d in enumerate(sdurs):
        rec = dict(recs[i], duration=d, num_samples=int(round(d * 16000))) if i < len(recs) else None
        srecs.append(rec)
        sups = [x for x in cuts[i]["supervisions"] if x["start"] + x["duration"] <= d]
        scuts.append({"id": f"rec{i}-0", "start": 0, "duration": d, "channel": 0, "supervisions": sups, "recording": rec, "type": "MonoCut"})
    for name, items in (("recordings_sincnet.jsonl.gz", srecs), ("cuts_sincnet.jsonl.gz", scuts)):
        with open(os.path.join(mdir, name), "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0) as f:
            for it in items:
                f.write((json.dumps(it) + "\n").encode())
    nfr_s = sum(math.ceil(rf.get_num_frames(16000 * d)) + 1 for d in sdurs)
    rows_s = (nfr_s + 292) // 293
    spreds = (torch.cumsum(torch.randn(rows_s * 293, generator=g), 0).reshape(rows_s, 293, 1) > 0).long()
    out["gnc_sincnet_preds"] = spreds.numpy().astype(np.uint8)
    sorig = {k: getattr(PS, k) for k in ("get_false_alarm", "get_missed_detection", "merge_intervals_with_buffer", "split_into_windows")}
    meta["get_new_cuts_sincnet"] = {"durations": sdurs}
    with tempfile.TemporaryDirectory() as td:
        torch.save(spreds, os.path.join(td, "preds.pt"))
        for tag, buffer, split in (("b0", 0, False), ("b1_split", 1, True)):
            log = {"fa": [], "md": [], "merged": [], "split": []}
            PS.get_false_alarm = lambda a, b, log=log: (log["fa"].append(float(sorig["get_false_alarm"](a, b))), sorig["get_false_alarm"](a, b))[1]
            PS.get_missed_detection = lambda a, b, log=log: (log["md"].append(float(sorig["get_missed_detection"](a, b))), sorig["get_missed_detection"](a, b))[1]
            PS.merge_intervals_with_buffer = lambda iv, d, b, log=log: (lambda r: (log["merged"].append([list(x) for x in r]), r)[1])(sorig["merge_intervals_with_buffer"](iv, d, b))
            PS.split_into_windows = lambda iv, window=10, log=log: (lambda r: (log["split"].append([list(x) for x in r]), r)[1])(sorig["split_into_windows"](iv, window=window))
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                PS.get_new_cuts("synthetic", "test", "preds.pt", os.path.join(mdir, "recordings_sincnet.jsonl.gz"),
                                os.path.join(mdir, "cuts_sincnet.jsonl.gz"), td, "unused.jsonl.gz", buffer=buffer, split=split)
            for k, v in sorig.items():
                setattr(PS, k, v)
            meta["get_new_cuts_sincnet"][tag] = {"buffer": buffer, "split": split, "fa": log["fa"], "md": log["md"],
                                                 "intervals": log["split"] if split else log["merged"], "report": buf.getvalue(),
                                                 "cuts": captured.pop("cuts")}

    # ---- a6 glue: binary_cross_entropy / interpolate (src/utils/loss.py:29-89), evaluated (and discarded) by _common_step
    from src.utils.loss import binary_cross_entropy
    bp = torch.rand(3, 40, 1, generator=g) * 0.98 + 0.01
    bt = (torch.rand(3, 40, generator=g) > 0.5).float()
    bw = torch.rand(3, 10, 1, generator=g)
    out["bce_pred"], out["bce_target"], out["bce_weight"] = bp.numpy(), bt.numpy(), bw.numpy()
    out["bce_plain"] = binary_cross_entropy(bp, bt, weight=None).numpy()
    out["bce_weighted"] = binary_cross_entropy(bp, bt, weight=bw).numpy()

    # ---- a12: load_config
    cfg = load_config()
    meta["config"] = {k: cfg[k] for k in ("seed", "device", "feature_extractor", "frame_shift", "model_name", "supported_models",
                                          "max_epochs", "learning_rate") if k in cfg}
    meta["config"]["model_dict"] = dict(cfg["model_dict"])
    for k in ("max_duration", "batch_size"):
        if k in cfg:
            meta["config"][k] = cfg[k]
    meta["torch_version"] = torch.__version__

    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True, default=str)
    print("reference fixtures written:", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "reference_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
