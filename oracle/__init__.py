"""TEST INFRASTRUCTURE: CPU oracle for the BBDuk k-mer path. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package (see oracle/bbduk_oracle.c)."""
