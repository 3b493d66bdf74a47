import json,sys
d=json.loads(sys.stdin.read())
print(d["ms"], [x for x in d["top_layers_us"] if x[0] in ("BN2","MaxPool4","BN5","Conv3","GlobAvg70")])
