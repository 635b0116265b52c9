import re, collections, sys
d = collections.OrderedDict()
for l in open(sys.argv[1]):
    m = re.match(r"(.*?)\s+kernel\s+([\d.]+) ms", l)
    if m and "saturated" not in l:
        d.setdefault(m.group(1).strip(), []).append(float(m.group(2)))
    elif "saturated" in l or "FAILED" in l or "Error" in l:
        print(l.strip())
for k, v in d.items():
    print("%-28s mean %.2f  n=%d  %s" % (k, sum(v) / len(v), len(v), " ".join("%.1f" % x for x in v)))
