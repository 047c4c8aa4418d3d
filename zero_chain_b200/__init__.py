"""zkb200 — B200-native Groth16 prover hot path for Zerochain's confidential-transfer circuit.

Host-side mirror of the reference's prover surface (bellman::groth16::{Parameters, Proof,
create_random_proof, create_proof}, bellman::multiexp::multiexp, bellman::domain::EvaluationDomain;
call sites core/proofs/src/confidential.rs:95-103,149) over the C ABI in include/zkb200.h.
The compute lives in csrc/ (hand-written sm_100a CUDA) behind libzkb200.so; this package is only
the ctypes binding used by tests and bench.py.  It fails loudly when the CUDA library is missing.
"""
