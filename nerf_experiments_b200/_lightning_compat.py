"""pytorch_lightning is optional: when it is installed the modules derive from the real
LightningModule (so `Trainer.fit` drives them exactly like the reference's); otherwise a minimal
stand-in keeps the same method surface so that plain training loops (nerf_experiments_b200.engine)
work.  Nothing on the compute path depends on Lightning."""
import torch as th
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        def __init__(self, *args, **kwargs):
            super().__init__()
            self.trainer = None
            self.automatic_optimization = True
            self._logged = {}

        def save_hyperparameters(self, *args, **kwargs):
            pass

        def log_dict(self, values, *args, **kwargs):
            self._logged.update(values)

        def log(self, name, value, *args, **kwargs):
            self._logged[name] = value

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            for b in self.buffers():
                return b.device
            return th.device("cpu")

        @property
        def dtype(self):
            return th.float32
