"""TEST TOOL: FASTA layout variants (CRLF, lower case, N / IUPAC / junk symbols, tabs and spaces in headers, no final newline, blank lines, empty records,
gzipped members) through tests/mock/agc-mock (product host objects + oracle-backed device ABI) and the reference binary: archives must be byte-identical."""
import os, subprocess, sys, tempfile, numpy as np, shutil, gzip
ROOT='/root/repo'
REF=ROOT+'/oracle/_ref/agc'; OUR=ROOT+'/tests/mock/agc-mock'
LET=b"ACGT"
def wrap(s, w, eol): return eol.join(s[i:i+w] for i in range(0,len(s),w))
def run(n_cases=60, seed=11, our=OUR, verbose=True):
  rng=np.random.default_rng(seed)
  def seq(n): return bytes(LET[i] for i in rng.integers(0,4,n))
  bad=0; n=0
  for case in range(n_cases):
      d=tempfile.mkdtemp(dir='/dev/shm')
      ref=seq(6000)
      files=[]
      for fi in range(3):
          eol=[b"\n",b"\r\n",b"\n"][case%3]
          recs=[]
          nrec=int(rng.integers(1,4))
          for r in range(nrec):
              s=bytearray(ref if r==0 else seq(int(rng.integers(50,3000))))
              for _ in range(20):
                  p=int(rng.integers(0,len(s))); s[p]=LET[int(rng.integers(0,4))]
              kind=case%10
              if kind==1: s=bytes(s).lower()
              if kind==2: s[100:110]=b"NNNNNNNNNN"
              if kind==3: s[200:203]=b"RYK"
              if kind==4: s[300:301]=b"*"
              body=wrap(bytes(s), int(rng.choice([60,70,80,10000])), eol)
              hdr=b">c%d_%d"%(fi,r) + (b"\tdesc with tab" if kind==5 else b" desc" if kind==6 else b"")
              rec=hdr+eol+body+(b"" if (kind==7 and r==nrec-1) else eol)
              if kind==8 and r==1: rec=hdr+eol+eol+body+eol+eol       # blank lines
              recs.append(rec)
          if case%10==9 and fi==2: recs.insert(1, b">empty_rec"+eol)     # empty record in the middle
          data=b"".join(recs)
          fn=os.path.join(d,"s%d.fa"%fi)
          if case%4==3 and fi==1:
              fn+=".gz"; open(fn,'wb').write(gzip.compress(data))
          else: open(fn,'wb').write(data)
          files.append(fn)
      flags=["-k","21","-s","1000","-b","3"] + (["-a"] if case%2 else [])
      outs=[]
      for exe,tag in ((REF,'r'),(our,'o')):
          out=os.path.join(d,tag+'.agc')
          subprocess.run([exe,"create"]+flags+["-o",out]+files,stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL)
          outs.append(open(out,'rb').read() if os.path.exists(out) else None)
      n+=1
      if outs[0]!=outs[1]: bad+=1; print("case",case,"differs", None if outs[0] is None else len(outs[0]), None if outs[1] is None else len(outs[1]))
      shutil.rmtree(d)
  if verbose: print(n,"cases, differing:",bad)
  return n, bad

if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 60)
