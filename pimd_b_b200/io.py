"""Text formats on either side of the hot path, kept byte-compatible with the reference.

* state dumps ``output/position_<b>.xyz``, ``velocity_<b>.dat``, ``force_<b>.dat``
  (reference src/states/position.cpp:33-54, velocity.cpp:33-53, force.cpp:33-53; ``{:^20.12e}`` columns, NDIM<3 padded
  with ``0.0`` columns);
* ``output/simulation.out`` (src/observables/observable.cpp:61-116; header ``{:^16s}``, values ``{:^16.8e}``);
* initial-state files read by ``initial_position = xyz(<fmt>)`` / ``initial_velocity = manual(<fmt>)``
  (src/common.cpp:49-136): positions in angstrom with exactly NDIM numeric columns, velocities in angstrom/ps after
  two ignored tokens, stored as ``m*v``.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Sequence

import numpy as np

from .config import convert_to_internal, convert_to_user


def _centered(text: str, width: int) -> str:
    """std::format's ``^`` alignment: extra padding goes to the right."""
    pad = width - len(text)
    if pad <= 0:
        return text
    left = pad // 2
    return " " * left + text + " " * (pad - left)


def fmt_e(value: float, width: int, prec: int) -> str:
    return _centered(f"{value:.{prec}e}", width)


# ------------------------------------------------------------------------------------------ state dumps
def format_frame(kind: str, step: int, arr: np.ndarray, ndim: int) -> str:
    """One frame of a per-bead dump; ``arr`` is [natoms][ndim] already in the user's unit."""
    n = arr.shape[0]
    pad = {1: " 0.0 0.0", 2: " 0.0", 3: ""}[ndim]
    lines = [f"{n}", f"Step {step}"]
    for i in range(n):
        head = "1" if kind == "position" else f"{i + 1} 1"
        cols = "".join(" " + fmt_e(float(arr[i, a]), 20, 12) for a in range(ndim))
        lines.append(head + cols + pad)
    return "\n".join(lines) + "\n"


class StateWriter:
    """PositionState / VelocityState / ForceState for all beads of this process (append mode, like the reference)."""

    FAMILY = {"position": "length", "velocity": "velocity", "force": "force"}
    FILE = {"position": "position_{}.xyz", "velocity": "velocity_{}.dat", "force": "force_{}.dat"}

    def __init__(self, kind: str, out_unit: str, freq: int, beads: Sequence[int], ndim: int, folder: str = "output"):
        try:
            self.factor = convert_to_user(self.FAMILY[kind], out_unit, 1.0)
        except ValueError:
            raise ValueError(f"Invalid output unit for {kind} state.")
        self.kind, self.freq, self.beads, self.ndim = kind, freq, list(beads), ndim
        os.makedirs(folder, exist_ok=True)
        self.files = [open(os.path.join(folder, self.FILE[kind].format(b)), "a") for b in self.beads]

    def output(self, step: int, arr: np.ndarray):
        """``arr``: [nbeads_local][natoms][ndim] in atomic units (velocity = momenta / mass done by the caller)."""
        if step % self.freq != 0:
            return
        for fh, slab in zip(self.files, arr):
            fh.write(format_frame(self.kind, step, slab * self.factor, self.ndim))

    def close(self):
        for fh in self.files:
            fh.close()


# ------------------------------------------------------------------------------------------ simulation.out
class ObservablesLogger:
    """ObservablesLogger (src/observables/observable.cpp:61-116)."""

    def __init__(self, columns: Iterable[str], folder: str = "output", filename: str = "simulation.out"):
        self.columns = list(columns)
        os.makedirs(folder, exist_ok=True)
        self.fh = open(os.path.join(folder, filename), "a")
        self.fh.write(_centered("step", 16) + "".join(" " + _centered(c, 16) for c in self.columns) + "\n")

    def log(self, step: int, values: Dict[str, float]):
        self.fh.write(fmt_e(float(step), 16, 8) + "".join(" " + fmt_e(values[c], 16, 8) for c in self.columns) + "\n")

    def close(self):
        self.fh.close()


def read_simulation_out(path: str) -> Dict[str, np.ndarray]:
    with open(path) as fh:
        header = fh.readline().split()
        data = np.loadtxt(fh, ndmin=2)
    return {name: data[:, i] for i, name in enumerate(header)}


# ------------------------------------------------------------------------------------------ initial-state files
def write_xyz_positions(path: str, x_au: np.ndarray):
    """[natoms][ndim] atomic units -> xyz in angstrom with 17 significant digits (NDIM numeric columns)."""
    f = convert_to_internal("length", "angstrom", 1.0)
    with open(path, "w") as fh:
        fh.write(f"{x_au.shape[0]}\n initial positions (angstrom)\n")
        for row in x_au:
            fh.write("He " + " ".join(f"{v / f:.17g}" for v in row) + "\n")


def write_manual_velocities(path: str, p_au: np.ndarray, mass: float):
    """[natoms][ndim] momenta (a.u.) -> LAMMPS-style velocity file in angstrom/ps (two leading tokens per row)."""
    f = convert_to_internal("velocity", "angstrom/ps", 1.0)
    with open(path, "w") as fh:
        fh.write(f"{p_au.shape[0]}\n initial velocities (angstrom/ps)\n")
        for i, row in enumerate(p_au):
            fh.write(f"{i + 1} 1 " + " ".join(f"{v / mass / f:.17g}" for v in row) + "\n")


def load_xyz_positions(path: str, natoms: int, ndim: int) -> np.ndarray:
    """loadTrajectories (src/common.cpp:49-97)."""
    if not os.path.exists(path):
        raise RuntimeError(f"Cannot open the xyz file named {path}.")
    with open(path) as fh:
        first = fh.readline().split()
        if not first or int(first[0]) != natoms:
            raise RuntimeError(f"The number of atoms in the xyz file ({path}) does not match the requested number of atoms.")
        fh.readline()
        tokens = fh.read().split()
    out = np.empty((natoms, ndim))
    f = convert_to_internal("length", "angstrom", 1.0)
    k = 0
    for i in range(natoms):
        k += 1  # symbol
        for a in range(ndim):
            out[i, a] = float(tokens[k]) * f
            k += 1
    return out


def load_manual_momenta(path: str, natoms: int, ndim: int, mass: float) -> np.ndarray:
    """loadMomenta (src/common.cpp:99-136)."""
    if not os.path.exists(path):
        raise RuntimeError(f"Cannot open the velocity file named {path}.")
    with open(path) as fh:
        first = fh.readline().split()
        if not first or int(first[0]) != natoms:
            raise RuntimeError(f"The number of atoms in the velocity file ({path}) does not match the requested number of atoms.")
        fh.readline()
        tokens = fh.read().split()
    out = np.empty((natoms, ndim))
    f = convert_to_internal("velocity", "angstrom/ps", 1.0)
    k = 0
    for i in range(natoms):
        k += 2
        for a in range(ndim):
            out[i, a] = mass * (float(tokens[k]) * f)
            k += 1
    return out


def read_dump_frames(path: str, ndim: int) -> List[np.ndarray]:
    """Frames of a position / velocity / force dump: the last three numeric columns of every row, cut to ndim."""
    frames = []
    with open(path) as fh:
        lines = fh.read().splitlines()
    i = 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        n = int(lines[i].split()[0])
        rows = [list(map(float, ln.split()[-3:])) for ln in lines[i + 2:i + 2 + n]]
        frames.append(np.asarray(rows)[:, :ndim])
        i += 2 + n
    return frames
