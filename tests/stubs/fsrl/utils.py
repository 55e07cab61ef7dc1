"""fsrl.utils loggers: the six methods the example scripts call."""


class DummyLogger:
    def __init__(self, *a, **k):
        self.rows, self.checkpoint_fn, self.saved = [], None, []

    def store(self, tab=None, **kw):
        self.rows.append(dict(kw))

    def save_config(self, cfg, verbose=False):
        self.config = cfg

    def setup_checkpoint_fn(self, fn):
        self.checkpoint_fn = fn

    def save_checkpoint(self, suffix=None):
        if self.checkpoint_fn is not None:
            self.saved.append((suffix, self.checkpoint_fn()))

    def write(self, step, display=False):
        pass

    write_without_reset = write


WandbLogger = TensorboardLogger = DummyLogger
