"""Import stub used ONLY by oracle/make_golden.py and the container-local pin tests.

The reference's BC path uses pytorch_lightning solely as a base class
(/root/reference/src/architectures/nets.py:6, src/models/imitation.py:27) plus
`self.log`, `self.lr_schedulers()`, `self.logger` and `self.current_epoch` inside
hooks. pytorch_lightning is not installed in this image, so this stub supplies
exactly that surface and nothing else. It is test infrastructure, not product code.
"""
import torch


class _Experiment:
    def __init__(self):
        self.scalars = []

    def add_scalars(self, tag, values, global_step=None):
        self.scalars.append((tag, {k: float(v) for k, v in values.items()}, global_step))


class _Logger:
    def __init__(self):
        self.experiment = _Experiment()


class LightningModule(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.logged = {}
        self.logger = _Logger()
        self.current_epoch = 0
        self._stub_schedulers = None

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    def lr_schedulers(self):
        return self._stub_schedulers


class Callback:
    pass
