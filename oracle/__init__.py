"""CPU oracle of the PackPPI-MSC sampling / PackPPI-Prox hot path.  TEST INFRASTRUCTURE, not product code:
only tests/, __graft_entry__.smoke() and bench.py's CPU legs (`cpu_baseline`, `--impl reference`) import it."""
