"""Imitation -- the behaviour-cloning training module, B200-native.

Hook-for-hook the contract of /root/reference/src/models/imitation.py:27-91
(`Imitation(hparams, net, data_loader)`; forward / training_step / validation_step /
*_epoch_end / *_dataloader / configure_optimizers / scale_image), so it drops in behind
train.py's behaviour_cloning block (train.py:93-129). What changes is underneath:
training_step runs one fused forward+CrossEntropy through the sm_100a kernels and returns an
autograd-connected scalar whose backward() runs the CUDA backward; configure_optimizers
returns the fused arena Adam instead of torch.optim.Adam (same hyper-parameters, same
state_dict keys), with the reference's MultiStepLR([20, 30], 0.1) on top.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.optim import lr_scheduler

try:
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    _HAVE_PL = True
except Exception:  # pragma: no cover - Lightning is not installed in the build image
    _Base = nn.Module
    _HAVE_PL = False

from carla_imitation_learning_b200.optim import FusedAdam


def lossCriterion(obj, inp, out):
    """The criterion of ImitationAux (/root/reference/src/models/imitation.py:11-24; name kept). In the reference every term but the
    last is commented out: loss = l3 = cross_entropy(inp[2], out[1][:, 1]) -- the autopilot-action logits (third output of the
    auxiliary net) against column 1 of the two-column label tensor. `out` is the module's `target = [x, y]`, so the fused kernels
    can compute exactly that from `out` alone: one launch chain for forward, CrossEntropy and (on backward) the gradients;
    `inp` (the separate forward the reference runs first) is not needed and may be None."""
    x, y = out
    if y.dim() != 2 or y.shape[1] < 2:
        raise IndexError("ImitationAux labels are (B, 2): [traffic-light status, autopilot action] (imitation.py:13-14)")
    fused = getattr(obj.net, "loss", None)
    if fused is None:
        raise TypeError("ImitationAux drives the B200 kernels through net.loss(x, y): pass a src.architectures.nets.ConvNet1")
    return fused(x, y[:, 1].contiguous())


class Imitation(_Base):
    def __init__(self, hparams, net, data_loader):
        super().__init__()
        self.h_params = hparams
        self.net = net
        self.data_loader = data_loader
        if not _HAVE_PL:
            self._logged = {}
            self._schedulers = None
            self.current_epoch = 0
            self.logger = None

    # -- what Lightning would provide; only defined when it is absent ------------------------
    if not _HAVE_PL:
        def log(self, name, value, **_kw):
            self._logged[name] = value

        def lr_schedulers(self):
            return self._schedulers

    # -- compute -------------------------------------------------------------------------------
    def forward(self, x):
        return self.net.forward(x)

    def _loss(self, x, y):
        fused = getattr(self.net, "loss", None)
        if fused is None:
            raise TypeError("Imitation drives the B200 kernels through net.loss(x, y): pass a src.architectures.nets.ConvNet1 "
                            "(there is no eager-PyTorch path in this module)")
        return fused(x, y)

    def training_step(self, batch, batch_idx):
        x, y = batch
        return self._loss(x, y)

    def validation_step(self, batch, batch_idx):
        x, y = batch
        with torch.no_grad():
            loss = self._loss(x, y)
        self.log('val_loss', loss)  # monitored by ModelCheckpoint (train.py:106-111)
        return loss

    # -- epoch hooks (imitation.py:57-71) -------------------------------------------------------
    def training_epoch_end(self, outputs) -> None:
        sch = self.lr_schedulers()
        if sch is not None:
            sch.step()
        # the one host synchronisation per epoch: a bounded device wait that expired, an out-of-range label or a peer that
        # never signalled would have left wrong numbers behind -- raise instead of training on
        eng = getattr(self.net, "_engine", None)
        if eng is not None:
            eng.check_device_errors()
            if getattr(eng, "peer", None) is not None:
                eng.peer.check()
        loss = torch.stack([o['loss'].detach() for o in outputs]).mean()   # stays on device until logged
        self._add_scalars({"train_loss": loss})

    def validation_epoch_end(self, outputs) -> None:
        self._add_scalars({"val_loss": torch.stack([o.detach() for o in outputs]).mean()})

    def _add_scalars(self, scalars) -> None:
        exp = getattr(getattr(self, "logger", None), "experiment", None)
        if exp is not None:
            exp.add_scalars("losses", scalars, global_step=self.current_epoch)

    # -- data (imitation.py:73-80) ---------------------------------------------------------------
    def train_dataloader(self):
        return self.data_loader['train_dataloader']

    def val_dataloader(self):
        return self.data_loader['val_dataloader']

    def test_dataloader(self):
        return self.data_loader['test_dataloader']

    # -- optimiser (imitation.py:82-87) ----------------------------------------------------------
    def configure_optimizers(self):
        params = list(self.parameters())
        if not all(hasattr(p, "_bc_arena") for p in params):
            raise TypeError("Imitation.configure_optimizers builds the fused arena Adam: every parameter must belong to a "
                            "src.architectures.nets.ConvNet1")
        optimizer = FusedAdam(params, lr=1e-3)         # LR is hard-coded in the reference (LEARNING_RATE is unused)
        scheduler = lr_scheduler.MultiStepLR(optimizer, milestones=[20, 30], gamma=0.1)
        if not _HAVE_PL:
            self._schedulers = scheduler
        return [optimizer], [scheduler]

    def scale_image(self, img):
        return (img + 1) / 2

    if not _HAVE_PL:
        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, hparams=None, net=None, data_loader=None, strict=True, **_kw):
            """pl.LightningModule.load_from_checkpoint as train.py:198-201 calls it (constructor kwargs passed in)."""
            from carla_imitation_learning_b200.trainer import load_checkpoint_into
            model = cls(hparams, net, data_loader)
            load_checkpoint_into(model, checkpoint_path, strict=strict)
            return model


class ImitationAux(Imitation):
    """The auxiliary-task module of /root/reference/src/models/imitation.py:94-159, hook for hook: `criterion = lossCriterion`,
    training_step / validation_step build `target = [x, y]` and return `criterion(self, output, target)`, validation does NOT
    log 'val_loss' (the reference's ImitationAux does not), epoch hooks / dataloaders / optimiser / scale_image as Imitation.
    The reference's lossCriterion keeps only the action term, so any ConvNet1 serves as `net` here: its logits ARE the third
    output the reference indexes (the image-reconstruction and traffic-light heads belong to CNNAuxNet, which the reference
    does not ship -- SURVEY 0.2). The separate `self.forward(x)` the reference runs before the criterion is skipped: the fused
    loss recomputes the forward inside the same launch chain."""

    def __init__(self, hparams, net, data_loader):
        super().__init__(hparams, net, data_loader)
        self.criterion = lossCriterion

    def training_step(self, batch, batch_idx):
        x, y = batch
        return self.criterion(self, None, [x, y])

    def validation_step(self, batch, batch_idx):
        x, y = batch
        with torch.no_grad():
            return self.criterion(self, None, [x, y])
