"""ORACLE package: CPU restatement of the reference's canonicalization hot path.

TEST INFRASTRUCTURE ONLY.  Importable from `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`; never imported by `equiadapt_b200/`.
"""
