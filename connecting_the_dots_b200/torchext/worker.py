"""Timing helpers the reference's torchext package exports (torchext/worker.py:19-78).  The training
Worker itself (host orchestration, checkpoints, matplotlib plots) is outside this framework's scope
(SURVEY.md section 2); `Worker` here only says so."""
import time


class StopWatch(object):
    """Named wall-clock timers: start(name) ... stop(name); get(name) -> accumulated seconds."""

    def __init__(self):
        self.timings = {}
        self.starts = {}

    def start(self, name):
        self.starts[name] = time.time()

    def stop(self, name):
        self.timings.setdefault(name, []).append(time.time() - self.starts.pop(name))

    def get(self, name=None, reduce=sum):
        if name is not None:
            return reduce(self.timings[name])
        return {k: reduce(v) for k, v in self.timings.items()}

    def __repr__(self):
        return ", ".join("%s=%.4f[s]" % (k, v) for k, v in self.get().items())

    __str__ = __repr__


class ETA(object):
    """Remaining-time estimate from the mean duration of the updates seen so far."""

    def __init__(self, length):
        self.length = length
        self.start_time = time.time()
        self.current_idx = 0
        self.current_time = time.time()

    def update(self, idx):
        self.current_idx = idx
        self.current_time = time.time()

    def get_elapsed_time(self):
        return self.current_time - self.start_time

    def get_item_time(self):
        return self.get_elapsed_time() / (self.current_idx + 1)

    def get_remaining_time(self):
        return self.get_item_time() * (self.length - self.current_idx + 1)

    @staticmethod
    def format_time(seconds):
        m, s = divmod(seconds, 60)
        h, m = divmod(m, 60)
        return "%02d:%02d:%02d" % (int(h), int(m), int(s))

    def get_elapsed_time_str(self):
        return self.format_time(self.get_elapsed_time())

    def get_remaining_time_str(self):
        return self.format_time(self.get_remaining_time())


class Worker(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("torchext.Worker (the reference's training loop, torchext/worker.py:46-528) is "
                                  "outside connecting_the_dots_b200's scope; only the op path is provided")
