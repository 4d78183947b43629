"""Seeded synthetic-input builders (the distributions of the reference's tests and benchmarks).
Neutral helper code: used by bench.py, tools/ and tests/, and re-exported by oracle/ for the tests.
Nothing here computes a reference result."""
