"""Expansion levels (reference: photon_weave/state/expansion_levels.py)."""

from enum import IntEnum


class ExpansionLevel(IntEnum):
    Label = 0
    Vector = 1
    Matrix = 2
