import random
P=0x18005
def crc16(data):
    c=0
    for b in data:
        c^=b<<8
        for _ in range(8):
            c=((c<<1)^0x8005)&0xFFFF if c&0x8000 else (c<<1)&0xFFFF
    return c
Q=0x8003
def fold16(v):  # v < 2^32 -> < 2^16 (lazy)
    h=v>>30; v=(v&0x3FFFFFFF)^h^(h<<2)
    h=v>>15; v=(v&0x7FFF)^h^(h<<1)
    return v
def fold15(v):
    while v>>15:
        h=v>>15; v=(v&0x7FFF)^h^(h<<1)
    return v
def horner_word(r,o):
    assert r < (1<<17)
    return fold16(((r<<4)^(r<<2)^o)&0xFFFFFFFFFFFF) if ((r<<4)^(r<<2)^o) < (1<<32) else None
def horner_byte(r,m):
    v=(r<<8)^m
    h=v>>15; v=(v&0x7FFF)^h^(h<<1)
    return v
def block_crc(words, tail_bytes):
    # words: list of 32-bit BE message words (leading zeros allowed), then tail bytes (1..4)
    N=len(words); nb=N//15; r_words=N-nb*15
    a=[0]*15
    idx=0
    for b in range(nb):
        a0=a[0]
        for i in range(15):
            o=words[idx]; idx+=1
            if i<14:
                a[i]=a[i]^a[i+1]^o
                if i==13: a[i]^=a0
            else:
                a[14]=a[14]^a0^o
    r=0
    if nb:
        for i in range(15):
            r=horner_word(r,a[i])
    for i in range(r_words):
        r=horner_word(r,words[idx]); idx+=1
    par=0
    for w in words: par^=w
    for m in tail_bytes:
        r=horner_byte(r,m); par^=m
    r=horner_byte(r,0); r=horner_byte(r,0)
    r=fold15(r)
    p=bin(par).count('1')&1
    t=(bin(r).count('1')&1)^p
    return r^(0x8003 if t else 0)
random.seed(1)
for trial in range(300):
    n=random.choice([680,1022,510,339,126,200,61,64,75])
    body=bytes(random.randrange(256) for _ in range(n))
    k=random.randint(1,4)          # tail bytes
    lead=(-(n-k))%4                # leading zeros to make word multiple
    pre=bytes(lead+4*random.randint(0,1))+body[:n-k]   # optionally one extra zero word
    assert len(pre)%4==0
    words=[int.from_bytes(pre[i:i+4],'big') for i in range(0,len(pre),4)]
    got=block_crc(words, body[n-k:])
    assert got==crc16(body),(n,k,hex(got),hex(crc16(body)))
print("ok")
