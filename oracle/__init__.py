"""oracle/ — TEST INFRASTRUCTURE.  CPU restatement of the reference's tile-raster path
(ifrit_oracle.c) and a ctypes driver for the reference's own CPU renderer built from
/root/reference (oracle/_ref/libifrit_ref.so).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package."""
