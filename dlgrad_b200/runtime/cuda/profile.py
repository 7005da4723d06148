"""CUDA-event timers on the compute stream (the stream every kernel of this backend launches on).

`KernelTimer` brackets individual launches of one kernel family inside a longer timed region (bench.py
uses it for the matmul launches of a train step: achieved FLOP/s = sum(flops) / sum(event time))."""
from __future__ import annotations

from dlgrad_b200.runtime.cuda import shim

lib = shim.lib


class Event:
    __slots__ = ("h",)

    def __init__(self) -> None:
        shim.init()
        self.h = lib.dlg_event_create()

    def record(self) -> "Event":
        shim.check(lib.dlg_event_record(self.h), "dlg_event_record")
        return self

    def sync(self) -> None:
        shim.check(lib.dlg_event_sync(self.h), "dlg_event_sync")

    def ms_since(self, start: "Event") -> float:
        return float(lib.dlg_event_elapsed_ms(start.h, self.h))


class KernelTimer:
    def __init__(self) -> None:
        self.pairs: list[tuple[Event, Event, float, tuple]] = []
        self._pool: list[Event] = []

    def _ev(self) -> Event:
        return self._pool.pop() if self._pool else Event()

    def start(self) -> Event:
        return self._ev().record()

    def stop(self, start: Event, work: float, tag: tuple = ()) -> None:
        self.pairs.append((start, self._ev().record(), work, tag))

    def collect(self) -> dict:
        """Call after a synchronize.  Returns total ms, total work and per-tag breakdown."""
        tot_ms, tot_work, by_tag = 0.0, 0.0, {}
        for a, b, work, tag in self.pairs:
            ms = b.ms_since(a)
            tot_ms += ms
            tot_work += work
            e = by_tag.setdefault(tag, [0, 0.0, 0.0])
            e[0] += 1
            e[1] += ms
            e[2] += work
            self._pool += [a, b]
        n = len(self.pairs)
        self.pairs = []
        return {"launches": n, "ms": tot_ms, "work": tot_work, "by_tag": by_tag}


matmul_timer: KernelTimer | None = None     # set by bench.py around the timed region
kernel_timer: KernelTimer | None = None     # every generated-kernel launch, by kernel family (bench.py --profile)
