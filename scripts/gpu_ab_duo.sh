#!/bin/bash
# A/B of the forward's two-images-per-tile mode (R <= 64): libhelio_v12.so = `make -C doodle_b200/csrc dbg DBGFLAGS=-DHELIO_FWD_DUO=0 DBGOUT=../libhelio_v12.so`;
# prints kernel time and an order-independent checksum of the image bits per build and shape (the two builds must agree bit for bit).
for rep in 1 2; do for lib in libhelio_v12.so libhelio_sm100.so; do for shape in "--N 5000 --R 64 --B 4096" "--N 500 --R 64 --B 16384" "--N 50 --R 64 --B 1025" "--N 300 --R 48 --B 777"; do echo -n "$lib $shape: "; HELIO_LIB_PATH=$PWD/doodle_b200/$lib timeout 120 python scripts/prof_splat.py --what fwd --impl 2 $shape --iters 3 2>&1 | tail -1; done; done; done
