"""Minimal stand-in for pytorch_lightning so the reference modules import on a box without it.
Only what the reference touches at import / construction / forward time."""
import torch.nn as nn


class LightningModule(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        self.trainer = None
        self.automatic_optimization = True

    def save_hyperparameters(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            import torch
            return torch.device("cpu")

    @property
    def dtype(self):
        import torch
        return torch.float32


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Callback:
    pass


class Trainer:
    def __init__(self, *a, **k):
        pass


def seed_everything(seed, workers=False):
    import torch
    torch.manual_seed(seed)
    return seed
