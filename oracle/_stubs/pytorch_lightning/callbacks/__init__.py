from .. import Callback


class LambdaCallback(Callback):
    def __init__(self, *a, **k):
        pass


class LearningRateMonitor(Callback):
    def __init__(self, *a, **k):
        pass


class ModelCheckpoint(Callback):
    def __init__(self, *a, **k):
        pass
