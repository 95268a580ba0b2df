"""Stopwatch that accumulates several start/pause intervals (host-side utility).

Interface of the reference's utils/timer.py (start / pause / stop / report, `duration`, `count`),
implemented on the monotonic high-resolution clock, plus a context-manager form:

    with Timer("epoch") as t:
        ...
    t.duration
"""
from time import perf_counter


class Timer(object):

    def __init__(self, task_name="UntitledTask"):
        self.task_name = task_name
        self._intervals = []     # closed intervals, seconds
        self._opened_at = None   # perf_counter() of the running interval, or None

    # -- state -----------------------------------------------------------------------------
    @property
    def is_timing(self):
        return self._opened_at is not None

    @property
    def check_point(self):
        return self._opened_at

    @property
    def duration(self):
        """total seconds over the closed intervals"""
        return float(sum(self._intervals))

    @property
    def count(self):
        """number of closed intervals"""
        return len(self._intervals)

    @property
    def mean(self):
        return self.duration / self.count if self._intervals else 0.0

    # -- control -----------------------------------------------------------------------------
    def start(self):
        if self._opened_at is None:          # a second start() while running is ignored
            self._opened_at = perf_counter()

    def pause(self):
        if self._opened_at is not None:      # a pause() while stopped is ignored
            self._intervals.append(perf_counter() - self._opened_at)
            self._opened_at = None

    def stop(self):
        self.pause()
        self.report()

    def reset(self):
        self._intervals, self._opened_at = [], None

    def report(self):
        print("[Timer] %s total: %.4f mean: %.4f count: %d"
              % (self.task_name, self.duration, self.mean, self.count))

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, exc_type, exc, tb):
        self.pause()
        return False
