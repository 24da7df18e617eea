"""CPU oracle (test infrastructure).  See topsy_oracle.py / splat_oracle.c headers.  Never imported by topsy_b200."""
