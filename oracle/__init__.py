"""TEST INFRASTRUCTURE: CPU oracle for the Groth16 prover hot path (see zk_oracle.c / pyref.py).
Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
