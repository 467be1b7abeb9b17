"""Wall-clock stopwatch with the reference's interface (cosypose/utils/timer.py:4-36):
start / pause / resume / stop, `stop()` returns the accumulated `datetime.timedelta`."""
import datetime


class Timer:
    def __init__(self):
        self.reset()

    def reset(self):
        self._t0 = None
        self.elapsed = datetime.timedelta(0)
        self.is_running = False

    def start(self):
        self.elapsed = datetime.timedelta(0)
        self._t0 = datetime.datetime.now()
        self.is_running = True
        return self

    def pause(self):
        if self.is_running:
            self.elapsed += datetime.datetime.now() - self._t0
            self.is_running = False

    def resume(self):
        if not self.is_running:
            self._t0 = datetime.datetime.now()
            self.is_running = True

    def stop(self):
        self.pause()
        elapsed = self.elapsed
        self.reset()
        return elapsed
