"""The reference's real train.py / train_ssl.py import and bind the CUDA-backed classes through the launcher
(eeg-gnn-ssl_b200/run.py), while the baseline models of train.py:20-22 still come from the reference.
Build-container test: it needs /root/reference (absent on the GPU box -> skipped there)."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DROPIN = os.path.join(ROOT, "eeg-gnn-ssl_b200")


@pytest.fixture(scope="module")
def stubs(tmp_path_factory):
    """the third-party packages the reference imports that this image lacks, as empty modules"""
    d = tmp_path_factory.mktemp("stubs")
    for n in ("h5py", "pyedflib", "dotted_dict", "tensorboardX"):
        (d / f"{n}.py").write_text({"dotted_dict": "DottedDict = dict\n",
                                    "tensorboardX": "SummaryWriter = object\n"}.get(n, ""))
    (d / "matplotlib").mkdir()
    (d / "matplotlib" / "__init__.py").write_text("from . import cm\n")
    (d / "matplotlib" / "cm.py").write_text("")
    return str(d)


def _check(script, stubs):
    env = dict(os.environ, PYTHONPATH=stubs)
    r = subprocess.run([sys.executable, os.path.join(DROPIN, "run.py"), "--dropin-check", os.path.join(REF, script)],
                       capture_output=True, text=True, env=env, cwd=REF, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("DROPIN ")][-1]
    return json.loads(line[7:])


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_train_py_binds_dropin_classes(stubs):
    got = _check("train.py", stubs)
    ours = os.path.join(DROPIN, "model", "model.py")
    assert got["DCRNNModel_classification"] == ours and got["DCRNNModel_nextTimePred"] == ours
    # baselines (train.py:20-22) are outside the path: they must still resolve, to the reference's own files
    assert got["LSTMModel"] == os.path.join(REF, "model", "lstm.py")
    assert got["CNN_LSTM"] == os.path.join(REF, "model", "cnnlstm.py")
    assert got["DenseCNN"] == os.path.join(REF, "model", "densecnn.py")


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_train_ssl_py_binds_dropin_classes(stubs):
    got = _check("train_ssl.py", stubs)
    assert got["DCRNNModel_nextTimePred"] == os.path.join(DROPIN, "model", "model.py")


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout")
def test_reference_train_fails_loudly_without_gpu(stubs):
    """the reference's train() over the drop-in classes on a CPU-only box: the product path refuses (no fallback)"""
    code = textwrap.dedent(f"""
        import sys, os, runpy, types, logging, tempfile, torch
        sys.path[:0] = [{DROPIN!r}, {REF!r}]
        os.environ["DCGRU_REFERENCE_ROOT"] = {REF!r}
        g = runpy.run_path({os.path.join(REF, 'train.py')!r}, run_name="x")
        a = types.SimpleNamespace(num_nodes=19, num_rnn_layers=2, rnn_units=64, input_dim=100, output_dim=100,
            max_diffusion_step=2, dcgru_activation="tanh", filter_type="laplacian", dropout=0.0, cl_decay_steps=3000,
            use_curriculum_learning=False, task="detection", model_name="dcrnn", metric_name="auroc",
            maximize_metric=True, lr_init=3e-4, l2_wd=5e-4, num_epochs=1, max_grad_norm=5.0, eval_every=9, patience=5)
        m = g["DCRNNModel_classification"](args=a, num_classes=1, device="cpu")
        class L(list):
            dataset = range(2)
        b = (torch.randn(2, 3, 19, 100), torch.ones(2), torch.full((2,), 3), [torch.eye(19).repeat(2, 1, 1)], None, None)
        class T:
            def add_scalar(self, *a): pass
        try:
            g["train"](m, {{"train": L([b]), "dev": L()}}, a, "cpu", tempfile.mkdtemp(), logging.getLogger("x"), T())
        except RuntimeError as e:
            print("REFUSED", str(e)[:80])
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env=dict(os.environ, PYTHONPATH=stubs), cwd=REF, timeout=300)
    if "cuda" in r.stdout.lower() and "REFUSED" not in r.stdout:
        pytest.skip("a GPU is present")
    assert "REFUSED" in r.stdout and "no CPU fallback" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
