#!/bin/bash
cd "$(dirname "$0")/.."
for pdl in 0 1; do echo "PDL=$pdl"; for n in 1024 98304 393216; do RESLIC_PDL=$pdl ./scratch/gcbench 24 $n 1 0 20; done; done
