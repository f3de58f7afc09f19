import collections, sys
ev=collections.defaultdict(list)
for ln in open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/trace.txt'):
    cta,role,tag,clk=map(int,ln.split()); ev[(cta,role)].append((tag,clk))
m=ev[(0,1)]; e=ev[(0,2)]; e7=ev[(0,0)]
starts=[ [c for t,c in e if t==i*10000+9000][0] for i in range(0,8)]
print("tile periods", [b-a for a,b in zip(starts,starts[1:])])
T=3; t0=starts[T]
allv=[(c-t0,'MMA',t) for t,c in m if t//10000==T]+[(c-t0,'EPI',t) for t,c in e if t//10000==T]+[(c-t0,'EP7',t) for t,c in e7 if t//10000==T]
names={0:'layer begin',1:'act ready',2:'acc_empty ok',3:'mma issued+commit',5:'acc_full ok',6:'epi chunk done',7:'act signalled',8:'16-col group done'}
for c,who,t in sorted(allv):
    l=(t%10000)//100; ch=(t%100)//10; ph=t%10
    if t%10000==9000: d='tile start'
    elif t%10000==9100: d='stage0 done'
    elif t%10000==9050: d='image constants loaded'
    else: d=f"L{l} c{ch} {names.get(ph,ph)}"
    print(f"{c:8d} {who} {d}")
