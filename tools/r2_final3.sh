# round-2 closing evidence, part 1: the whole GPU suite on the final code
O=gpurun_out/r2t; mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/gputest.txt; cat $O/gputest.txt
