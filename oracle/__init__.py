"""CPU checker of the B200 CLIP hot path — TEST INFRASTRUCTURE ONLY.

`clip_oracle` restates the reference algorithm in plain fp32 / fp64 torch ops (each function cites the reference file:line it
follows); `make_golden` pins it to outputs of the unmodified reference (`tests/golden/`); `ref_loader` imports the reference
in the build container.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU arm may import this package; the
product path (`understanding_clip_ood_b200`) never does.
"""
