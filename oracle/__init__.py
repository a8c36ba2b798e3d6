"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's GP hot path.
Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
