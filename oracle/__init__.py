"""CPU oracle for the CDC decode hot path -- TEST INFRASTRUCTURE, NOT THE PRODUCT.

The upstream reference (/root/reference) ships no code: README.md is 0 bytes and
.gitignore:1-27 is a stock Python template.  Per BASELINE.json `north_star` the
parity source is therefore this small PyTorch-eager fp32 restatement of a
CDC-style codec (Yang & Mandt, "Lossy Image Compression with Conditional
Diffusion Models"), specified in SURVEY.md Appendix A.

PARITY UNPINNED BY THE REFERENCE: the reference holds no golden vectors, tests
or runnable implementation for this path.  What pins the
oracle instead are the analytic known-answer tests in tests/test_oracle_*.py
(DDIM identity, schedule indices, half-to-even rounding, CDF monotonicity,
threshold ties), cross-checks of every composite block against
torch.nn.functional primitives, and the committed golden vectors under
tests/golden/ produced by oracle/make_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product
(conditional-diffusion-model-for-compression_b200/) never does.
"""
from .config import CDCConfig  # noqa: F401
