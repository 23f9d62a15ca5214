cd tools/micro
for mode in 0 1 2 3 4 5 6; do for W in 64 128 256; do ./membench $W $mode 4; done; done
for mode in 1 2; do for W in 64 128; do ./membench $W $mode 2; ./membench $W $mode 8; ./membench $W $mode 4 8320; done; done
