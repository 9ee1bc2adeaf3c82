"""ImitationBranched -- EXTENSION (no counterpart in the reference): the Imitation module contract
(/root/reference/src/models/imitation.py:27-91: training_step / validation_step / *_epoch_end / *_dataloader /
configure_optimizers) for the command-conditioned ConvNet1Branched. A batch is (x, command, target): target = class ids
for branch_loss 'ce', (B, n_out) float32 (steer, throttle, brake) for 'l1' / 'mse'. Specification: oracle/ext_oracle.py."""
from __future__ import annotations

import torch
from torch.optim import lr_scheduler

from .imitation import Imitation, _HAVE_PL


class ImitationBranched(Imitation):
    def forward(self, x, command):
        return self.net.forward(x, command)

    def training_step(self, batch, batch_idx):
        x, command, target = batch
        return self.net.loss(x, command, target)

    def validation_step(self, batch, batch_idx):
        x, command, target = batch
        with torch.no_grad():
            loss = self.net.loss(x, command, target)
        self.log('val_loss', loss)
        return loss

    def configure_optimizers(self):
        optimizer = self.net.configure_optimizer(lr=1e-3)          # same Adam hyper-parameters as imitation.py:83
        scheduler = lr_scheduler.MultiStepLR(optimizer, milestones=[20, 30], gamma=0.1)
        if not _HAVE_PL:
            self._schedulers = scheduler
        return [optimizer], [scheduler]
