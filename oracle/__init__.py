"""TEST INFRASTRUCTURE. CPU restatements of the reference algorithms (the oracle) and loaders for
the unmodified reference kernels built into oracle/_ref/.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / reference legs may import this package; the product
(animatablegaussians_b200) never does."""
