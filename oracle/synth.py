"""Seeded synthetic inputs live in vittf_b200/synth.py (they are generators, not hot-path
algorithms, and bench.py's product arm needs them without importing oracle/); re-exported here for
the oracle-side scripts."""
from vittf_b200.synth import annotations, class_features, ct_volume, shell_labels, torus_volume  # noqa: F401
