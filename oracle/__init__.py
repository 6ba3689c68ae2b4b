"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the SP-GAN hot path (kNN graph + EdgeConv generator, PointNet
critic, WGAN-GP penalty, train step).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product
(`sp-gan_b200/`) never does and fails loudly when its CUDA library is missing.

Parity status: the reference (liruihui/SP-GAN) owns no tests or golden vectors for
this path ("parity unpinned" by reference-owned fixtures).  The oracle is pinned
instead against outputs of the unmodified reference modules run on CPU in the build
container: tests/golden/make_golden.py generated tests/golden/*.npz, and
tests/test_oracle_*.py check the oracle against them.
"""
