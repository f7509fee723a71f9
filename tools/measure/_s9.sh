#!/bin/bash
cd /root/repo
timeout 600 python tools/measure/chunk_bits.py 2>&1 | tail -12
