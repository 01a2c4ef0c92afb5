"""oracle/ — TEST INFRASTRUCTURE, not product code.

A CPU restatement (torch-CPU / numpy, fp32 or fp64) of the reference hot path
(TUM-DAML/gemnet_pytorch @ a0164f7): GemNet.forward and the host index builder.  Every
function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import anything from here, and only as the checker / the timed CPU baseline.  The
product package `gemnet_pytorch_b200` never imports `oracle` and has no CPU fallback.

Pinning: `oracle/make_golden.py` imports the UNMODIFIED reference (through `oracle/ref_shims.py`)
in the build container, runs it on COLL test molecules, and writes `tests/golden/*.npz`.
`tests/test_oracle_golden.py` checks this restatement against those fixtures (runs anywhere) and
`tests/test_oracle_vs_reference.py` checks it against the live reference with the shipped
pretrained checkpoints whenever `/root/reference` is present.
"""
