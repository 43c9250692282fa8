#!/bin/bash
timeout 600 python tools/two_stream_bench.py --batch 8 --streams 1
timeout 600 python tools/two_stream_bench.py --batch 8 --streams 2
timeout 600 python tools/two_stream_bench.py --batch 16 --streams 2
timeout 600 python tools/two_stream_bench.py --batch 16 --streams 1
