import sys
sys.path.insert(0,'/root/repo')
import leafline_b200 as ll
eng=ll.Engine(0)
root=ll.new_world()
for W in (8,4,2):
    parts=[eng.perft(root,7,shard=(r,W))[0] for r in range(W)]
    tot=sum(parts)
    print(W, tot, "max/mean %.4f"%(max(parts)/(tot/W)), [round(100*p/tot,2) for p in parts])
kiwi=ll.reconstruct("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -")
parts=[eng.perft(kiwi,6,shard=(r,8))[0] for r in range(8)]
tot=sum(parts); print("kiwi d6", tot, "max/mean %.4f"%(max(parts)/(tot/8)))
