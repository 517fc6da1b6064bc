"""Golden fixtures of the reference's OWN training loops and fine-tune path (acceptance harness).

Run in the build container only (needs /root/reference):

    python tests/golden/make_train_golden.py

It imports the UNMODIFIED ``train.py`` / ``train_ssl.py`` of the reference (SURVEY Appendix B stub recipe) and
calls their ``train()`` functions -- ``train.py:197-329``, ``train_ssl.py:101-230`` -- over the reference's own
model classes on CPU with a synthetic in-memory loader, recording the loss the loop logs every step
(``tbx.add_scalar('train/Loss', ...)``) and the weights it ends with.  The GPU box has no /root/reference, so
``tests/test_gpu_trainloop.py`` replays the same batches through the drop-in modules with the loop body restated
and compares against these files.

trainloop_cls.npz   BASELINE config 1 (detection, distance graph, T=12, K=2, H=64, L=2, B=4), 3 steps of train.train()
trainloop_ssl.npz   README SSL setting (12 s in / 12 s out, L=3, H=64, K=2, distance graph), B=3, 3 steps of
                    train_ssl.train() incl. the inverse-scaled loss (utils.py:460-495; note the call passes
                    loss_fn="MAE", which the helper's `== 'mae'` test sends to masked_mse_loss: the loss the
                    reference really trains on is the masked RMSE)
pretrained_dist12.npz   pretrained/pretrained_distance_graph_12s.pth.tar loaded strict into
                    DCRNNModel_nextTimePred: predictions, loss and every gradient; then
                    utils.build_finetune_model (utils.py:166-176) into a 2-layer detection model: logits, gradients
"""
import logging
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, Args, import_reference, make_supports, np32, save, structured_clip  # noqa: E402


class Loader(list):
    """what train() touches of a DataLoader: iteration and len(loader.dataset)"""

    @property
    def dataset(self):
        return range(sum(b[0].shape[0] for b in self))


class Tbx:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, val, step):
        self.scalars.setdefault(tag, []).append(float(val))


def train_args(**kw):
    a = Args(task="detection", model_name="dcrnn", metric_name="auroc", maximize_metric=True, lr_init=3e-4,
             l2_wd=5e-4, num_epochs=1, max_grad_norm=5.0, eval_every=1000, patience=5, num_classes=1,
             rand_seed=123)
    a.__dict__.update(kw)
    return a


def state_arrays(model, prefix):
    """state_dict as arrays; the aliases of the tied decoder cell (decoding_cells.2.* == decoding_cells.1.*,
    model/model.py:142-143) are left out to keep the files small -- the test re-creates them"""
    return {f"{prefix}:{k}": np32(v) for k, v in model.state_dict().items() if ".decoding_cells.2." not in k}


def cls_loop(mm, ru, du, rtrain):
    args = train_args()
    ru.seed_torch(seed=args.rand_seed)
    model = mm.DCRNNModel_classification(args=args, num_classes=1, device="cpu")
    gen = torch.Generator().manual_seed(77)
    batches = []
    for _ in range(3):
        raw = torch.stack([structured_clip(gen, 12, 19, 100) for _ in range(4)])
        x = (raw - 3.924) / 1.560
        sup, _ = make_supports(ru, du, mm, "laplacian", raw.numpy())
        y = (torch.rand(4, generator=gen) > 0.5).float()
        sl = torch.full((4,), 12, dtype=torch.long)
        batches.append((x, y, sl, sup, None, None))
    # initial weights are NOT stored: utils.seed_torch(123) + construction is reproduced seed-for-seed by the drop-in
    arrays = {"init_checksum": np.array([float(sum(v.double().sum() for v in model.state_dict().values()))])}
    for i, (x, y, sl, sup, _, _) in enumerate(batches):
        arrays[f"x{i}"], arrays[f"y{i}"], arrays[f"sl{i}"] = np32(x), np32(y), sl.numpy()
        arrays[f"sup{i}"] = np32(sup[0])
    tbx = Tbx()
    with tempfile.TemporaryDirectory() as d:
        rtrain.train(model, {"train": Loader([(x, y, sl, list(sup), a, b) for x, y, sl, sup, a, b in batches]),
                             "dev": Loader()}, args, "cpu", d, logging.getLogger("golden"), tbx)
    arrays["loss"] = np.array(tbx.scalars["train/Loss"], dtype=np.float64)
    arrays.update(state_arrays(model, "final"))
    meta = dict(kind="trainloop_cls", batch=4, T=12, steps=3, lr_init=args.lr_init, l2_wd=args.l2_wd,
                max_grad_norm=args.max_grad_norm, **{k: getattr(args, k) for k in Args().__dict__})
    save("trainloop_cls", meta, arrays)


def ssl_loop(mm, ru, du, rssl):
    args = train_args(num_rnn_layers=3, metric_name="loss", maximize_metric=False)
    ru.seed_torch(seed=args.rand_seed)
    model = mm.DCRNNModel_nextTimePred(args=args, device="cpu")
    gen = torch.Generator().manual_seed(78)
    batches = []
    for _ in range(3):
        raw = torch.stack([structured_clip(gen, 24, 19, 100) for _ in range(3)])
        xy = (raw - 3.924) / 1.560
        x, y = xy[:, :12].contiguous(), xy[:, 12:].contiguous()
        sup, _ = make_supports(ru, du, mm, "laplacian", raw[:, :12].numpy())
        batches.append((x, y, None, sup, None, None))
    arrays = {"init_checksum": np.array([float(sum(v.double().sum() for v in model.state_dict().values()))])}
    for i, (x, y, _, sup, _, _) in enumerate(batches):
        arrays[f"x{i}"], arrays[f"y{i}"], arrays[f"sup{i}"] = np32(x), np32(y), np32(sup[0])
    scaler = ru.StandardScaler(mean=np.float64(3.924), std=np.float64(1.560))
    tbx = Tbx()
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self            # model/model.py:336 hard-codes .cuda()
    try:
        with tempfile.TemporaryDirectory() as d:
            rssl.train(model, {"train": Loader([(x, y, s, list(sup), a, b) for x, y, s, sup, a, b in batches]),
                               "dev": Loader()}, args, "cpu", d, logging.getLogger("golden"), tbx, scaler=scaler)
    finally:
        torch.Tensor.cuda = orig_cuda
    arrays["loss"] = np.array(tbx.scalars["train/MAE Loss"], dtype=np.float64)
    arrays.update(state_arrays(model, "final"))
    meta = dict(kind="trainloop_ssl", batch=3, T=12, To=12, steps=3, lr_init=args.lr_init, l2_wd=args.l2_wd,
                max_grad_norm=args.max_grad_norm, scaler_mean=3.924, scaler_std=1.560,
                **{k: getattr(args, k) for k in Args().__dict__})
    save("trainloop_ssl", meta, arrays)


def pretrained_case(mm, ru, du):
    ckpt = os.path.join(REF, "pretrained", "pretrained_distance_graph_12s.pth.tar")
    args3 = train_args(num_rnn_layers=3)
    torch.manual_seed(5)
    pre = mm.DCRNNModel_nextTimePred(args=args3, device="cpu")
    pre = ru.load_model_checkpoint(ckpt, pre)                  # strict=True inside (utils.py:158)
    gen = torch.Generator().manual_seed(79)
    raw = torch.stack([structured_clip(gen, 24, 19, 100) for _ in range(2)])
    xy = (raw - 3.924) / 1.560
    x, y = xy[:, :12].contiguous(), xy[:, 12:].contiguous()
    sup, _ = make_supports(ru, du, mm, "laplacian", raw[:, :12].numpy())
    pre.train()
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        pred = pre(x, y, sup, batches_seen=0)
    finally:
        torch.Tensor.cuda = orig_cuda
    loss = ru.compute_regression_loss(y_true=y, y_predicted=pred, loss_fn="MAE",
                                      standard_scaler=ru.StandardScaler(mean=np.float64(3.924), std=np.float64(1.560)), device="cpu")
    loss.backward()
    arrays = {"x": np32(x), "y": np32(y), "support0": np32(sup[0]), "pred": np32(pred), "loss": np32(loss)}
    arrays.update(state_arrays(pre, "ckpt"))
    for n_, p_ in pre.named_parameters():
        arrays["sslgrad:" + n_] = np32(p_.grad)
    # fine-tune transplant (train.py:133-148): a fresh 2-layer detection model takes the encoder cells' sub-modules
    args2 = train_args(num_rnn_layers=2)
    torch.manual_seed(6)
    new = mm.DCRNNModel_classification(args=args2, num_classes=1, device="cpu")
    arrays["newinit:fc.weight"], arrays["newinit:fc.bias"] = np32(new.fc.weight), np32(new.fc.bias)
    new = ru.build_finetune_model(model_new=new, model_pretrained=pre, num_rnn_layers=args2.num_rnn_layers)
    new.train()
    new.zero_grad()
    sl = torch.tensor([12, 9], dtype=torch.long)
    xz = x.clone()
    xz[1, 9:] = 0
    logits = new(xz, sl, sup)
    yb = torch.tensor([1.0, 0.0])
    l2 = torch.nn.BCEWithLogitsLoss()(logits.view(-1), yb)
    l2.backward()
    arrays.update({"ft_x": np32(xz), "ft_sl": sl.numpy(), "ft_y": np32(yb), "ft_logits": np32(logits),
                   "ft_loss": np32(l2)})
    for n_, p_ in new.named_parameters():
        arrays["ftgrad:" + n_] = np32(p_.grad)
    arrays["ft_state_keys"] = np.array(list(new.state_dict().keys()))
    meta = dict(kind="pretrained", checkpoint="pretrained/pretrained_distance_graph_12s.pth.tar", batch=2, T=12, To=12,
                **{k: getattr(args3, k) for k in Args().__dict__})
    save("pretrained_dist12", meta, arrays)


def main():
    mm, ru, du = import_reference()
    sys.modules["tensorboardX"].SummaryWriter = Tbx
    import train as rtrain              # noqa: E402  (the reference's train.py, unmodified)
    import train_ssl as rssl            # noqa: E402
    torch.set_num_threads(8)
    logging.basicConfig(level=logging.WARNING)
    cls_loop(mm, ru, du, rtrain)
    ssl_loop(mm, ru, du, rssl)
    pretrained_case(mm, ru, du)


if __name__ == "__main__":
    main()
