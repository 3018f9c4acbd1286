"""CPU oracle (test infrastructure).  Importable only from tests/, smoke() and bench.py's CPU legs."""
