"""ORACLE — test infrastructure only.

A CPU restatement of the reference's COLLECT -> CLUSTER path.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package; nothing under `svim_b200/` does.
"""
