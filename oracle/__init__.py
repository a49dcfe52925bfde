"""CPU oracle (test infrastructure only). See oracle/dgl_oracle.py."""
