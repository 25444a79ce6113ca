import sys, os
sys.path.insert(0, '/root/repo')
from svim_b200 import synth, io as sio, _lib
batch, _g, _ = synth.make_config("config2", 0.02, with_genome=False)
p = "/tmp/c2.bam"
sio.write_bam_native(p, batch, threads=8)
ctx = _lib.Context(device=0)
for i in range(2):
    st = {}
    rb = sio.decode_bam_resident(p, ctx, st)
    print(st)
