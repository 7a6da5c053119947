import sys; sys.path.insert(0,'.')
import barretenberg_b200 as bb
lib=bb.default_library()
for mode,name in ((3,'fq_mul 64 warps/SM, 2 chains'),(5,'fr_mul 16 warps/SM, 1 chain'),(6,'fr_mul 16 warps/SM, 2 chains'),(7,'fr_mul 16 warps/SM, 4 chains')):
    ops,ms=lib.microbench(mode,512)
    print('%-34s %.3e mul/s  (%.2f ms)'%(name,ops,ms))
