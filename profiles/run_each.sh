#!/bin/bash
# runs each collected test id in its own process (a faulting kernel poisons the CUDA context)
while read id; do
  timeout 120 python -m pytest "$id" -q -m gpu 2>&1 | tail -1 | sed "s|^|$id  |"
done
