#!/bin/bash
( time timeout 1500 python -m pytest tests -q -m gpu -x -k "every_compiled_fast_variant" 2>&1 | tail -6 ) 2>&1 | tail -10
