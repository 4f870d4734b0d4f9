#!/bin/bash
# Session-4 experiment: chunk plan of the host entry point on config 3.
EVEREST_HOST_FIRST_DIV=8 python tools/probe_host_chunks.py d8
EVEREST_HOST_FIRST_DIV=4 python tools/probe_host_chunks.py d4
EVEREST_HOST_FIRST_DIV=16 python tools/probe_host_chunks.py d16
EVEREST_HOST_FIRST_DIV=16 EVEREST_HOST_THREE=1 python tools/probe_host_chunks.py d16x3
EVEREST_HOST_FIRST_DIV=32 EVEREST_HOST_THREE=1 python tools/probe_host_chunks.py d32x3
EVEREST_HOST_FIRST_DIV=1000000 python tools/probe_host_chunks.py minb
