"""Cutoff sweep over lists of tensors (mirror of the reference's ``evaluation/benchmark.py`` core loop)."""
