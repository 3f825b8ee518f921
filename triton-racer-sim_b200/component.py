"""The plugin base type — same constructor, attributes and hooks as the reference's
TritonRacerSim/components/component.py:3-28, so `Car.addComponent` (core/car.py:16-25) and `DataPool`
(core/datapool.py:8-28) drive these classes unchanged.  When the reference package is importable the batched
components subclass *its* Component (car.py:17 asserts issubclass); otherwise this identical declaration."""
from abc import ABC

try:  # pragma: no cover - only when the reference is on sys.path
    from TritonRacerSim.components.component import Component as _RefComponent
except Exception:  # noqa: BLE001
    _RefComponent = None


class _Component(ABC):
    def __init__(self, inputs=[], outputs=[], threaded=False):
        """Names of input and output values are strings (e.g. 'cam/img')."""
        self.step_inputs = inputs.copy()
        self.step_outputs = outputs.copy()
        self.threaded = threaded

    def onStart(self):
        """Called right before the main loop begins."""

    def step(self, *args):
        """Behaviour in the main loop: takes the values of step_inputs, returns a tuple for step_outputs."""

    def thread_step(self):
        """Behaviour in the component's own thread (threaded components only)."""

    def onShutdown(self):
        """Shutdown."""

    def getName(self):
        return 'Generic Component'


Component = _RefComponent if _RefComponent is not None else _Component
