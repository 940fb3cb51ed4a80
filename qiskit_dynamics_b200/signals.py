"""Signals: time-dependent scalar coefficients  s(t) = Re[f(t) exp(i(2 pi nu t + phi))].

Host-side mirror of the reference's ``qiskit_dynamics.signals`` (signals/signals.py): same class
names, constructor arguments and evaluation semantics.  Signals are *coefficient sources*: the
solvers evaluate a whole ``SignalList`` on the stage-time grid in one vectorised NumPy call
(:meth:`SignalList.table`) and ship the resulting ``(T, K)`` table to HBM, where the fused CUDA
kernels consume it.  Envelopes may be arbitrary Python callables, which
is why this layer is NumPy on the host and not a kernel.

Bin-edge semantics of :class:`DiscreteSignal` follow NumPy float floor-division exactly
(signals/signals.py:302-311).
"""

from __future__ import annotations

import itertools
import operator as _operator
from typing import Callable, List, Optional, Sequence, Union

from itertools import chain as _chain
from operator import attrgetter as _attrgetter

import numpy as np

from .exceptions import QiskitError

_TWO_PI_J = 2j * np.pi

# Mutation epoch: bumped by every in-place change of a signal (carrier_freq / phase setters, add_samples).  Models cache the
# compiled device program of their signals together with the epoch it was compiled at and recompile when it has moved, so
# that the device route sees in-place edits exactly as the host route -- and the reference, which re-reads the signal
# objects on every evaluation (signals/signals.py:148-155) -- does.
_epoch = 0


def _bump_epoch() -> None:
    global _epoch
    _epoch += 1


def mutation_epoch() -> int:
    return _epoch


def _scalar_like(x) -> bool:
    return isinstance(x, (int, float, complex, np.number)) or (hasattr(x, "ndim") and np.ndim(x) == 0)


class Signal:
    """Envelope times complex carrier; ``signal(t)`` is the real part (signals/signals.py:34-155)."""

    def __init__(self, envelope: Union[Callable, complex, float, int], carrier_freq=0.0, phase=0.0,
                 name: Optional[str] = None):
        self._name = name
        self._is_constant = False
        self._const_envelope = None  # numeric envelope (device-evaluable, see compile_signal_program)
        if callable(envelope):
            self._envelope = envelope
        else:
            const = np.asarray(envelope)
            if const.ndim == 0:
                self._const_envelope = complex(const)
            if np.all(np.asarray(carrier_freq) == 0.0):
                self._is_constant = True
            self._envelope = lambda t, _c=const: _c * np.ones_like(t)
        self.carrier_freq = carrier_freq
        self.phase = phase

    # -- parameters ---------------------------------------------------------------------------
    @property
    def name(self):
        return self._name

    @property
    def is_constant(self) -> bool:
        return self._is_constant

    @property
    def carrier_freq(self):
        return self._carrier_freq

    @carrier_freq.setter
    def carrier_freq(self, value):
        _bump_epoch()
        self._carrier_freq = np.asarray(value)
        self._carrier_arg = _TWO_PI_J * self._carrier_freq

    @property
    def phase(self):
        return self._phase

    @phase.setter
    def phase(self, value):
        _bump_epoch()
        self._phase = np.asarray(value)
        self._phase_arg = 1j * self._phase

    # -- evaluation ---------------------------------------------------------------------------
    def envelope(self, t):
        return self._envelope(t)

    def complex_value(self, t):
        return self.envelope(t) * np.exp(self._carrier_arg * t + self._phase_arg)

    def __call__(self, t):
        return np.real(self.complex_value(t))

    def conjugate(self) -> "Signal":
        env = self._envelope
        return Signal(lambda t: np.conjugate(env(t)), -self.carrier_freq, -self.phase)

    # -- algebra ------------------------------------------------------------------------------
    def __add__(self, other):
        return signal_add(self, other)

    def __radd__(self, other):
        return signal_add(other, self)

    def __mul__(self, other):
        return signal_multiply(self, other)

    def __rmul__(self, other):
        return signal_multiply(other, self)

    def __neg__(self):
        return signal_multiply(-1, self)

    def __sub__(self, other):
        return signal_add(self, -other)

    def __rsub__(self, other):
        return signal_add(other, -self)

    def __str__(self):
        if self._name is not None:
            return str(self._name)
        if self.is_constant:
            return f"Constant({self(0.0)})"
        return f"Signal(carrier_freq={self.carrier_freq}, phase={self.phase})"


class DiscreteSignal(Signal):
    """Piecewise-constant envelope given by samples of width ``dt`` (signals/signals.py:257-313)."""

    def __init__(self, dt: float, samples, start_time: float = 0.0, carrier_freq=0.0, phase=0.0, name=None):
        self._dt = dt
        self._start_time = start_time
        self._set_samples(np.asarray(samples))
        Signal.__init__(self, self._lookup, carrier_freq=carrier_freq, phase=phase, name=name)

    def _set_samples(self, samples: np.ndarray):
        if samples.shape[0] == 0:
            pad = np.zeros((1,) + samples.shape[1:], dtype=samples.dtype if samples.dtype != object else float)
        else:
            pad = np.zeros_like(samples[:1])
        # one trailing zero: index -1 (before start) and index N (after end) both read it
        self._padded = np.concatenate([samples, pad], axis=0)

    def _lookup(self, t):
        t = np.asarray(t)
        # NumPy float floor-division on purpose: this *is* the reference's bin-edge rule
        idx = np.clip(np.array((t - self._start_time) // self._dt, dtype=int), -1, len(self.samples))
        return self._padded[idx]

    @classmethod
    def from_Signal(cls, signal: Signal, dt: float, n_samples: int, start_time: float = 0.0,
                    sample_carrier: bool = False) -> "DiscreteSignal":
        mid = start_time + (np.arange(n_samples) + 0.5) * dt
        if sample_carrier:
            return DiscreteSignal(dt, signal(mid), start_time=start_time, carrier_freq=0.0, phase=signal.phase,
                                  name=signal.name)
        return DiscreteSignal(dt, signal.envelope(mid), start_time=start_time, carrier_freq=signal.carrier_freq,
                              phase=signal.phase, name=signal.name)

    @property
    def duration(self) -> int:
        return len(self.samples)

    @property
    def dt(self) -> float:
        return self._dt

    @property
    def samples(self) -> np.ndarray:
        return self._padded[:-1]

    @property
    def start_time(self) -> float:
        return self._start_time

    def conjugate(self):
        return self.__class__(dt=self._dt, samples=np.conjugate(self.samples), start_time=self._start_time,
                              carrier_freq=-self.carrier_freq, phase=-self.phase)

    def add_samples(self, start_sample: int, samples):
        _bump_epoch()
        samples = np.asarray(samples)
        if len(samples) < 1:
            return
        current = self.samples
        if start_sample < len(current):
            raise QiskitError("Samples can only be added afer the last sample.")
        gap = start_sample - len(current)
        if gap > 0:
            current = np.append(current, np.zeros(gap, dtype=samples.dtype))
        self._set_samples(np.append(current, samples))

    def __str__(self):
        if self._name is not None:
            return str(self._name)
        return f"DiscreteSignal(dt={self.dt}, carrier_freq={self.carrier_freq}, phase={self.phase})"


class SignalCollection:
    """List-like container with NumPy-style subscripting (signals/signals.py:450-503)."""

    def __init__(self, signal_list: List[Signal]):
        self._is_constant = False
        self._components = signal_list

    @property
    def components(self) -> List[Signal]:
        return self._components

    def __len__(self):
        return len(self.components)

    def __iter__(self):
        return iter(self.components)

    def __getitem__(self, idx):
        if not isinstance(idx, slice) and np.ndim(idx) > 0:
            picked = [self.components[int(i)] for i in idx]
            return self.__class__(picked) if len(picked) != 1 else picked[0]
        sub = _operator.itemgetter(idx)(self.components)
        return self.__class__(sub) if isinstance(sub, list) else sub

    def conjugate(self):
        return self.__class__([s.conjugate() for s in self.components])


class SignalSum(SignalCollection, Signal):
    """Sum of signals; evaluates all terms in one vectorised pass (signals/signals.py:505-609)."""

    def __init__(self, *signals, name: Optional[str] = None):
        self._name = name
        terms: List[Signal] = []
        for sig in signals:
            if isinstance(sig, list):
                sig = SignalSum(*sig)
            if isinstance(sig, SignalSum):
                terms.extend(sig.components)
            elif isinstance(sig, Signal):
                terms.append(sig)
            elif _scalar_like(sig):
                terms.append(Signal(sig))
            else:
                raise QiskitError("Components of a SignalSum must be instances of a Signal subclass or a scalar.")
        SignalCollection.__init__(self, terms)
        Signal.__init__(self, self._stacked_envelope, carrier_freq=[s.carrier_freq for s in terms],
                        phase=[s.phase for s in terms], name=name)

    def _stacked_envelope(self, t):
        return np.moveaxis(np.asarray([s.envelope(t) for s in self.components]), 0, -1)

    def complex_value(self, t):
        carriers = np.exp(np.expand_dims(t, -1) * self._carrier_arg + self._phase_arg)
        return np.sum(self.envelope(t) * carriers, axis=-1)

    def flatten(self) -> Signal:
        if len(self) == 0:
            return Signal(0.0)
        if len(self) == 1:
            return self.components[0]
        mean_freq = np.sum(self.carrier_freq) / len(self)
        rel_arg = self._carrier_arg - _TWO_PI_J * mean_freq

        def merged(t):
            carriers = np.exp(np.expand_dims(t, -1) * rel_arg + self._phase_arg)
            return np.sum(self.envelope(t) * carriers, axis=-1)

        return Signal(envelope=merged, carrier_freq=mean_freq, name=str(self))

    def __str__(self):
        if self._name is not None:
            return str(self._name)
        if len(self) == 0:
            return "SignalSum()"
        return " + ".join(str(s) for s in self.components)


class DiscreteSignalSum(DiscreteSignal, SignalSum):
    """Sum of piecewise-constant signals sharing dt / start / duration; samples are (N, terms)
    (signals/signals.py:612-777)."""

    def __init__(self, dt: float, samples, start_time: float = 0.0, carrier_freq=None, phase=None, name=None):
        samples = np.asarray(samples)
        nterms = samples.shape[-1]
        carrier_freq = np.zeros(nterms) if carrier_freq is None else carrier_freq
        phase = np.zeros(nterms) if phase is None else phase
        DiscreteSignal.__init__(self, dt=dt, samples=samples, start_time=start_time, carrier_freq=carrier_freq,
                                phase=phase, name=name)
        self._components = [
            DiscreteSignal(dt=self.dt, samples=col, start_time=self.start_time, carrier_freq=f, phase=p)
            for col, f, p in zip(self.samples.transpose(), np.atleast_1d(carrier_freq), np.atleast_1d(phase))
        ]

    @classmethod
    def from_SignalSum(cls, signal_sum: SignalSum, dt: float, n_samples: int, start_time: float = 0.0,
                       sample_carrier: bool = False) -> "DiscreteSignalSum":
        mid = start_time + (np.arange(n_samples) + 0.5) * dt
        freq = signal_sum.carrier_freq
        samples = signal_sum.envelope(mid)
        if sample_carrier:
            samples = samples * np.exp(np.expand_dims(mid, -1) * signal_sum._carrier_arg)
            freq = 0.0 * freq
        return DiscreteSignalSum(dt, samples, start_time=start_time, carrier_freq=freq, phase=signal_sum.phase,
                                 name=signal_sum.name)

    def __getitem__(self, idx):
        if isinstance(idx, int) and idx >= len(self):
            raise IndexError(f"index out of range for DiscreteSignalSum of length {len(self)}")
        samples = self.samples[:, idx]
        freqs = np.atleast_1d(self.carrier_freq[idx])
        phases = np.atleast_1d(self.phase[idx])
        if samples.ndim == 1:
            return DiscreteSignal(dt=self.dt, samples=samples, start_time=self.start_time, carrier_freq=freqs[0],
                                  phase=phases[0])
        if samples.shape[1] == 1:
            return DiscreteSignal(dt=self.dt, samples=samples[:, 0], start_time=self.start_time,
                                  carrier_freq=freqs[0], phase=phases[0])
        return DiscreteSignalSum(dt=self.dt, samples=samples, start_time=self.start_time, carrier_freq=freqs,
                                 phase=phases)

    def __str__(self):
        if self._name is not None:
            return str(self._name)
        if len(self) == 0:
            return "DiscreteSignalSum()"
        return " + ".join(str(s) for s in self.components)


class SignalList(SignalCollection):
    """K signals evaluated together: ``siglist(t)`` -> (K,) or (T, K) (signals/signals.py:780-835)."""

    def __init__(self, signal_list: Sequence[Union[Signal, float, complex]]):
        super().__init__([to_SignalSum(s) for s in signal_list])

    def complex_value(self, t):
        return np.moveaxis(np.asarray([s.complex_value(t) for s in self.components]), 0, -1)

    def __call__(self, t):
        return np.moveaxis(np.asarray([s(t) for s in self.components]), 0, -1)

    def table(self, times) -> np.ndarray:
        """Coefficient table for a grid of T times: float64 (T, K), C-contiguous -- the device format."""
        times = np.asarray(times, dtype=float)
        return np.ascontiguousarray(self(times), dtype=np.float64).reshape(times.shape[0], len(self))

    def flatten(self) -> "SignalList":
        return SignalList([s.flatten() if isinstance(s, SignalSum) else s for s in self.components])

    @property
    def drift(self) -> np.ndarray:
        """Sum of the constant terms of each entry."""
        out = []
        for entry in self.components:
            total = 0.0
            for term in (entry if isinstance(entry, SignalSum) else SignalSum(entry)):
                if term.is_constant:
                    total = total + term(0.0)
            out.append(total)
        return np.asarray(out)


# ---------------------------------------------------------------------------------------------
# device form: flattened term arrays for qdb_signal_table_f64
# ---------------------------------------------------------------------------------------------


class SignalProgram:
    """Flattened description of one SignalList -- or of B structurally identical SignalLists, one per
    state column -- that the device evaluates on a time grid (``csrc/signals.cu``).

    Terms are the elementary components of every channel in channel order: piecewise-constant
    (:class:`DiscreteSignal`) or constant-envelope (:class:`Signal` with a numeric envelope) signals.  Arrays
    (host, NumPy): ``chan`` (nterms,) int32; ``samp_len`` (nterms,) int32 (-1 = constant envelope);
    ``samp_off`` (nterms,) int64; ``dt``, ``t0``, ``freq``, ``phase`` (nterms,) float64 -- or (nterms, B)
    when they differ between columns; ``samples`` complex128, (S,) shared or (B, S) per column.
    """

    def __init__(self, num_channels, columns, chan, samp_len, samp_off, dt, t0, freq, phase, samples):
        self.num_channels, self.columns = num_channels, columns
        self.chan, self.samp_len, self.samp_off = chan, samp_len, samp_off
        self.dt, self.t0, self.freq, self.phase = dt, t0, freq, phase
        self.samples = samples
        self._dev = None

    @property
    def params_per_column(self) -> bool:
        return self.freq.ndim == 2

    @property
    def samples_per_column(self) -> bool:
        return self.samples.ndim == 2

    def to_device(self, device):
        import torch

        if self._dev is None or self._dev["device"] != device:
            t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(device)  # noqa: E731
            self._dev = {
                "device": device,
                "terms": {"chan": t(self.chan, np.int32), "samp_off": t(self.samp_off, np.int64),
                          "samp_len": t(self.samp_len, np.int32), "dt": t(self.dt, np.float64), "t0": t(self.t0, np.float64),
                          "freq": t(self.freq, np.float64), "phase": t(self.phase, np.float64)},
                "samples": t(self.samples if self.samples.size else np.zeros(1, complex), np.complex128),
            }
        return self._dev

    def table(self, times, device):
        """(T, K) -- or (T, K, B) for a per-column program -- float64 device tensor; a Python float gives T = 1."""
        import torch

        from . import _abi

        d = self.to_device(device)
        if np.ndim(times) == 0 and not isinstance(times, torch.Tensor):
            times_dev = float(times)  # one time: passed by value
        else:
            times_dev = times if isinstance(times, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(times, dtype=np.float64)).to(device)
        stride = self.samples.shape[1] if self.samples_per_column else 0
        return _abi.signal_table(self.num_channels, d["terms"], d["samples"], times_dev, B=self.columns, col_stride=stride,
                                 params_per_col=self.params_per_column)


def _term_descr(term):
    """(samp_len, samples, dt, t0, freq, phase) of an elementary device-evaluable signal, else None."""
    if isinstance(term, SignalSum):
        return None
    freq, phase = np.asarray(term.carrier_freq), np.asarray(term.phase)
    if freq.ndim != 0 or phase.ndim != 0 or np.iscomplexobj(freq) or np.iscomplexobj(phase):
        return None
    if type(term) is DiscreteSignal:
        samples = np.asarray(term.samples)
        if samples.ndim != 1 or samples.dtype == object:
            return None
        return (samples.shape[0], samples.astype(complex), float(term.dt), float(term.start_time), float(freq), float(phase))
    if type(term) is Signal and term._const_envelope is not None:
        return (-1, np.asarray([term._const_envelope], dtype=complex), 1.0, 0.0, float(freq), float(phase))
    return None


def _compile_discrete_lists(lists) -> Optional[SignalProgram]:
    """compile_signal_program for the shape large pulse sweeps have -- B plain lists of K DiscreteSignals with scalar
    carrier and phase -- written against the objects' fields with preallocated arrays (a quarter of the host time of
    the general route); None when the input is anything else.  Sample counts may differ between simulations (pulses of
    different durations): channel j is padded with zeros to its longest simulation, which is what a DiscreteSignal
    evaluates to past its end -- the padding to a common duration of the reference's batched schedule path
    (solvers/solver_classes.py:607-660)."""
    B, first = len(lists), lists[0]
    K = len(first)
    if K == 0 or any(type(x) is not DiscreteSignal for x in first):
        return None
    if any(x._padded.ndim != 1 or x._padded.dtype == object for x in first):
        return None
    # C-level passes over the B K objects (itertools / operator): a Python-level generator per object costs three times as much
    if set(map(len, lists)) != {K}:
        return None
    flat = list(_chain.from_iterable(lists))
    if set(map(type, flat)) != {DiscreteSignal}:
        return None
    pads = list(map(_attrgetter("_padded"), flat))
    if set(map(_attrgetter("ndim"), pads)) != {1}:
        return None
    plen = np.fromiter(map(len, pads), dtype=np.int64, count=B * K).reshape(B, K)
    lens = [int(v) for v in plen.max(axis=0) - 1]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    try:
        params = np.fromiter(_chain.from_iterable(map(_attrgetter("_dt", "_start_time", "_carrier_freq", "_phase"), flat)),
                             dtype=float, count=4 * B * K).reshape(B, K, 4)
        if bool(np.all(plen == plen[:1])):
            # every simulation has the same sample counts (the usual sweep: amplitudes, frequencies, phases vary): one
            # concatenation of all sample arrays instead of B K slice assignments -- a third of the host time
            # The concatenation IS the sample store: channel j of a simulation starts at the sum of the PADDED lengths before
            # it (every DiscreteSignal carries one trailing zero, which stays in the store and is never indexed) -- no
            # second pass over the 16 B x samples x simulations
            samples = np.concatenate(pads).astype(complex, copy=False).reshape(B, int(plen[0].sum()))
            offs = np.concatenate([[0], np.cumsum(plen[0])]).astype(np.int64)
        else:  # ragged sample counts: the same layout (padded length of the longest simulation per channel), filled one by one
            offs = np.concatenate([[0], np.cumsum(plen.max(axis=0))]).astype(np.int64)
            samples = np.zeros((B, int(offs[-1])), dtype=complex)
            it = iter(pads)
            for b in range(B):
                row = samples[b]
                for j in range(K):
                    pad = next(it)
                    row[offs[j]:offs[j] + pad.shape[0]] = pad
    except (TypeError, ValueError):  # array-valued or complex carrier / phase, object samples: the general route decides
        return None
    # (a sweep almost always differs already in its second simulation: look there before comparing everything)
    shared_params = B == 1 or (bool(np.array_equal(params[1], params[0])) and bool(np.all(params == params[:1])))
    shared_samples = B == 1 or (bool(np.array_equal(samples[1], samples[0])) and bool(np.all(samples == samples[:1])))
    pick = (lambda k: params[0, :, k].copy()) if shared_params else (lambda k: np.ascontiguousarray(params[:, :, k].T))
    return SignalProgram(K, B, np.arange(K, dtype=np.int32), np.asarray(lens, dtype=np.int32), offs[:-1].copy(),
                         pick(0), pick(1), pick(2), pick(3), samples[0].copy() if shared_samples else samples)


def compile_signal_program(signal_lists) -> Optional[SignalProgram]:
    """Device program of one SignalList, or of a list of SignalLists (sweep mode: one per column) that share
    their structure (same channels, same kind and sample count of every term).  None when a term is an
    arbitrary Python envelope -- those stay on the host path (:meth:`SignalList.table`)."""
    single = isinstance(signal_lists, SignalList)
    lists = [signal_lists] if single else list(signal_lists)
    if not lists:
        return None
    if not single and isinstance(lists[0], (list, tuple)):
        fast = _compile_discrete_lists(lists)
        if fast is not None:
            return fast
    per_col = []
    for sl in lists:
        descr = []
        # a plain list of signals is read as it is: wrapping every simulation of a large sweep into a SignalList
        # (one DiscreteSignalSum per channel) costs more host time than the solve itself
        entries = sl.components if isinstance(sl, SignalList) else sl
        for j, entry in enumerate(entries):
            terms = entry.components if isinstance(entry, SignalSum) else (entry,)
            for term in terms:
                d = _term_descr(term) if isinstance(term, Signal) else None
                if d is None:
                    return None
                descr.append((j,) + d)
        per_col.append(descr)
    ref = per_col[0]
    sig = [(d[0], d[1]) for d in ref]
    if any([(d[0], d[1]) for d in descr] != sig for descr in per_col[1:]):
        return None
    nterms, B = len(ref), len(lists)
    chan = np.asarray([d[0] for d in ref], dtype=np.int32)
    samp_len = np.asarray([d[1] for d in ref], dtype=np.int32)
    counts = np.asarray([len(d[2]) for d in ref], dtype=np.int64)
    samp_off = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64) if nterms else np.zeros(0, np.int64)
    params = np.asarray([[d[3:7] for d in descr] for descr in per_col], dtype=float).reshape(B, nterms, 4)
    flat = [np.concatenate([d[2] for d in descr]) if nterms else np.zeros(0, complex) for descr in per_col]
    shared_params = bool(np.all(params == params[:1]))
    shared_samples = all(np.array_equal(f, flat[0]) for f in flat[1:])
    pick = (lambda k: params[0, :, k].copy()) if shared_params else (lambda k: np.ascontiguousarray(params[:, :, k].T))
    samples = flat[0] if shared_samples else np.stack(flat)
    num_channels = len(lists[0]) if isinstance(lists[0], (SignalList, list, tuple)) else len(list(lists[0]))
    return SignalProgram(num_channels, 0 if single else B, chan, samp_len, samp_off, pick(0), pick(1), pick(2), pick(3), samples)


# ---------------------------------------------------------------------------------------------
# algebra (signals/signals.py:838-1121)
# ---------------------------------------------------------------------------------------------


def to_SignalSum(sig) -> SignalSum:
    if isinstance(sig, SignalSum):
        return sig
    if isinstance(sig, DiscreteSignal):
        samples = sig.samples
        cols = samples.reshape(1, 0) if samples.shape == (0,) else samples.reshape(-1, 1)
        return DiscreteSignalSum(dt=sig.dt, samples=cols, start_time=sig.start_time,
                                 carrier_freq=np.asarray([sig.carrier_freq]), phase=np.asarray([sig.phase]))
    if isinstance(sig, Signal):
        return SignalSum(sig)
    if not isinstance(sig, list) and _scalar_like(sig):
        return SignalSum(Signal(sig))
    raise QiskitError("Input type incompatible with SignalSum.")


def _same_grid(a: DiscreteSignal, b: DiscreteSignal) -> bool:
    return a.dt == b.dt and a.start_time == b.start_time and a.duration == b.duration


def signal_add(sig1, sig2) -> SignalSum:
    try:
        a, b = to_SignalSum(sig1), to_SignalSum(sig2)
    except QiskitError as err:
        raise QiskitError("Only a number or a Signal instance can be added to a Signal.") from err
    if isinstance(a, DiscreteSignalSum) and isinstance(b, DiscreteSignalSum) and _same_grid(a, b):
        return DiscreteSignalSum(dt=a.dt, samples=np.append(a.samples, b.samples, axis=1), start_time=a.start_time,
                                 carrier_freq=np.append(a.carrier_freq, b.carrier_freq),
                                 phase=np.append(a.phase, b.phase))
    return SignalSum(*(a.components + b.components))


def _rank(sig) -> int:
    """Order used to put the 'more special' operand first: constant < DiscreteSignal < Signal <
    SignalSum < DiscreteSignalSum (signals/signals.py:1052-1079)."""
    if sig.is_constant:
        return 0
    if isinstance(sig, DiscreteSignalSum):
        return 4
    if isinstance(sig, DiscreteSignal):
        return 1
    if isinstance(sig, SignalSum):
        return 3
    return 2


def sort_signals(sig1: Signal, sig2: Signal):
    return (sig1, sig2) if _rank(sig1) <= _rank(sig2) else (sig2, sig1)


def base_signal_multiply(sig1: Signal, sig2: Signal) -> Signal:
    """Product of two elementary signals via  Re[a]Re[b] = Re[ab]/2 + Re[a conj(b)]/2."""
    a, b = sort_signals(sig1, sig2)
    if a.is_constant and b.is_constant:
        return Signal(a(0.0) * b(0.0))
    if a.is_constant and type(b) is DiscreteSignal:
        return DiscreteSignal(dt=b.dt, samples=a(0.0) * b.samples, start_time=b.start_time,
                              carrier_freq=b.carrier_freq, phase=b.phase)
    if a.is_constant and type(b) is Signal:
        c = a(0.0)
        return Signal(envelope=lambda t: c * b.envelope(t), carrier_freq=b.carrier_freq, phase=b.phase)
    if type(a) is DiscreteSignal and type(b) is DiscreteSignal and _same_grid(a, b):
        plus = DiscreteSignal(dt=b.dt, samples=0.5 * a.samples * b.samples, start_time=b.start_time,
                              carrier_freq=a.carrier_freq + b.carrier_freq, phase=a.phase + b.phase)
        minus = DiscreteSignal(dt=b.dt, samples=0.5 * a.samples * np.conjugate(b.samples), start_time=b.start_time,
                               carrier_freq=a.carrier_freq - b.carrier_freq, phase=a.phase - b.phase)
        return plus + minus
    plus = Signal(lambda t: 0.5 * a.envelope(t) * b.envelope(t), a.carrier_freq + b.carrier_freq, a.phase + b.phase)
    minus = Signal(lambda t: 0.5 * a.envelope(t) * np.conjugate(b.envelope(t)), a.carrier_freq - b.carrier_freq,
                   a.phase - b.phase)
    return plus + minus


def signal_multiply(sig1, sig2) -> SignalSum:
    try:
        a, b = to_SignalSum(sig1), to_SignalSum(sig2)
    except QiskitError as err:
        raise QiskitError("Only a number or a Signal instance can multiply a Signal.") from err
    a, b = sort_signals(a, b)
    if len(a) == 1 and a[0].is_constant and isinstance(b, DiscreteSignalSum):
        return DiscreteSignalSum(dt=b.dt, samples=a(0.0) * b.samples, start_time=b.start_time,
                                 carrier_freq=b.carrier_freq, phase=b.phase)
    if isinstance(a, DiscreteSignalSum) and isinstance(b, DiscreteSignalSum) and _same_grid(a, b):
        n = a.samples.shape[0]
        pair = (a.samples[:, :, None] * b.samples[:, None, :]).reshape(n, -1)
        pair_conj = (a.samples[:, :, None] * b.samples[:, None, :].conj()).reshape(n, -1)
        return DiscreteSignalSum(
            dt=a.dt, samples=np.append(0.5 * pair, 0.5 * pair_conj, axis=1), start_time=a.start_time,
            carrier_freq=np.append(np.add.outer(a.carrier_freq, b.carrier_freq).reshape(-1),
                                   np.subtract.outer(a.carrier_freq, b.carrier_freq).reshape(-1)),
            phase=np.append(np.add.outer(a.phase, b.phase).reshape(-1), np.subtract.outer(a.phase, b.phase).reshape(-1)))
    product = SignalSum()
    for x, y in itertools.product(a.components, b.components):
        product = product + base_signal_multiply(x, y)
    return product
