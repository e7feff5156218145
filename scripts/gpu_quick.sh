#!/bin/bash
# quick session on the GPU box: the whole GPU test suite
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8
