"""Error type of the reference surface.  The reference raises ``qiskit.QiskitError``
(e.g. models/operator_collections.py:119-122); qiskit is not a dependency here, so a local
class of the same name is used (re-exported from qiskit when that package is importable)."""

try:  # pragma: no cover - qiskit is absent in the build image
    from qiskit import QiskitError  # type: ignore
except Exception:  # noqa: BLE001

    class QiskitError(Exception):
        """Base error of the qiskit-dynamics surface."""

        def __init__(self, *message):
            super().__init__(" ".join(str(m) for m in message))
            self.message = " ".join(str(m) for m in message)

        def __str__(self):
            return repr(self.message)
