import re,collections,sys
d=collections.OrderedDict()
for l in open(sys.argv[1]):
    m=re.match(r"(\S+)\s+kernel\s+([\d.]+) ms.*chk (\w+)",l)
    if m: d.setdefault(m.group(1),[]).append((float(m.group(2)),m.group(3)))
for k,v in d.items():
    t=[x[0] for x in v]
    print(k, "n %d mean %.2f median %.2f max %.1f  n>32ms %d"%(len(t),sum(t)/len(t),sorted(t)[len(t)//2],max(t),sum(1 for x in t if x>32)), "chk", v[0][1], v[-1][1])
