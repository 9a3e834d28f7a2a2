#!/usr/bin/env python
"""Plan-9 (Go assembler) AMD64 -> GNU as (AT&T) transliterator.

TEST INFRASTRUCTURE ONLY (see oracle/minlz_oracle.h): this script is the recipe
that makes the reference's own hot loops runnable in an image without a Go
toolchain.  It reads the reference's generated assembly where it lies
(`/root/reference/asm_amd64.s`, never copied into this repo), rewrites every
instruction 1:1 into AT&T syntax and writes the result to `oracle/_ref/`
(git-ignored).  `oracle/ref_shim.c` supplies the Go ABI0 call frame.

What is translated (everything the file uses, ~65 mnemonics):
  * operand order is already src, dst in both syntaxes; only CMPx swaps
    (Plan 9 `CMPQ a, b` sets flags for a - b  ==  AT&T `cmpq b, a`);
  * registers take their width from the mnemonic suffix (AX -> %rax/%eax/%ax/%al);
    address registers are always 64 bit;
  * `name+off(FP)` -> off+frame+8(%rsp) (locals below, return address, then the
    ABI0 argument block); `off(SP)` is the hardware SP after `sub $frame, %rsp`;
  * TEXT -> global symbol `p9_<name>` + frame allocation; RET -> frame release + ret;
  * PCALIGN $n -> .p2align log2(n).
Anything it does not know raises: a silent mistranslation would poison the
baseline.

usage: python oracle/p9_to_gas.py /root/reference/asm_amd64.s oracle/_ref/minlz_asm_amd64.S
"""
import re
import sys

GP = {
    "AX": ("rax", "eax", "ax", "al"), "BX": ("rbx", "ebx", "bx", "bl"),
    "CX": ("rcx", "ecx", "cx", "cl"), "DX": ("rdx", "edx", "dx", "dl"),
    "SI": ("rsi", "esi", "si", "sil"), "DI": ("rdi", "edi", "di", "dil"),
    "BP": ("rbp", "ebp", "bp", "bpl"),
}
for _i in range(8, 16):
    GP["R%d" % _i] = ("r%d" % _i, "r%dd" % _i, "r%dw" % _i, "r%db" % _i)
BYTE_ALIAS = {"AL": "al", "BL": "bl", "CL": "cl", "DL": "dl"}
SIZE_IDX = {"q": 0, "l": 1, "w": 2, "b": 3}

JCC = {"JMP": "jmp", "JB": "jb", "JBE": "jbe", "JAE": "jae", "JA": "ja", "JNE": "jne", "JNZ": "jnz",
       "JZ": "jz", "JEQ": "je", "JE": "je", "JNA": "jna", "JL": "jl", "JG": "jg", "JC": "jc",
       "JNC": "jnc", "JLT": "jl"}
CMOV_CC = {"LT": "l", "GE": "ge", "LE": "le", "EQ": "e"}

# mnemonic -> (AT&T mnemonic, operand sizes); sizes: q/l/w/b for GP registers, x for xmm
SIMPLE = {}
for _base in ("MOV", "ADD", "SUB", "XOR", "OR", "AND", "TEST", "SHL", "SHR", "SAR", "INC", "DEC", "CMP"):
    for _s in "QLWB":
        SIMPLE[_base + _s] = (_base.lower() + _s.lower(), _s.lower())
SIMPLE.update({
    "LEAQ": ("leaq", "q"), "LEAL": ("leal", "l"), "IMULQ": ("imulq", "q"), "TZCNTQ": ("tzcntq", "q"),
    "BTL": ("btl", "l"),
})
ZX = {"MOVWLZX": ("movzwl", "w", "l"), "MOVWQZX": ("movzwq", "w", "q"), "MOVBQZX": ("movzbq", "b", "q"),
      "MOVBLZX": ("movzbl", "b", "l"), "MOVLQZX": ("movl", "l", "l")}
SSE = {"MOVOU": "movdqu", "MOVOA": "movdqa", "PXOR": "pxor"}


class Fn:
    frame = 0
    name = ""


def local_label(name, fn):
    return ".L%s.%s" % (fn.name, name)


def reg(name, size):
    if name in BYTE_ALIAS:
        return "%" + BYTE_ALIAS[name]
    if name in GP:
        return "%" + GP[name][SIZE_IDX[size]]
    m = re.fullmatch(r"X(\d+)", name)
    if m:
        return "%xmm" + m.group(1)
    raise ValueError("unknown register " + name)


MEM_RE = re.compile(r"^(?P<disp>-?(?:0x[0-9a-fA-F]+|\d+))?\((?P<base>[A-Z0-9]+)\)(?:\((?P<idx>[A-Z0-9]+)\*(?P<sc>[1248])\))?$")
FP_RE = re.compile(r"^[A-Za-z_][A-Za-z0-9_]*\+(\d+)\(FP\)$")


def operand(op, size, fn):
    op = op.strip()
    if op.startswith("$"):
        return op
    m = FP_RE.match(op)
    if m:
        return "%d(%%rsp)" % (int(m.group(1)) + fn.frame + 8)
    m = MEM_RE.match(op)
    if m:
        disp = m.group("disp") or ""
        base = m.group("base")
        if base == "SP":
            if m.group("idx"):
                raise ValueError("indexed SP operand " + op)
            d = int(disp, 0) if disp else 0
            if not 0 <= d < fn.frame:
                raise ValueError("SP operand outside the frame: " + op)
            return "%d(%%rsp)" % d
        s = disp + "(" + reg(base, "q")
        if m.group("idx"):
            s += "," + reg(m.group("idx"), "q") + "," + m.group("sc")
        return s + ")"
    if re.fullmatch(r"[A-Z][A-Z0-9]*", op):
        return reg(op, size)
    raise ValueError("unknown operand " + op)


def split_ops(s):
    return [x.strip() for x in s.split(",")] if s.strip() else []


def translate_line(line, fn, out):
    raw = line.rstrip("\n")
    s = raw.strip()
    if not s or s.startswith("//"):
        return
    if s.startswith("#include"):
        return
    m = re.match(r"^TEXT\s+·(\w+)\(SB\),\s*(?:NOSPLIT,\s*)?\$(\d+)-(\d+)$", s)
    if m:
        fn.name, fn.frame = m.group(1), int(m.group(2))
        out.append("\n\t.text\n\t.p2align 5\n\t.globl p9_%s\n\t.type p9_%s,@function\np9_%s:" % ((fn.name,) * 3))
        if fn.frame:
            out.append("\tsubq $%d, %%rsp" % fn.frame)
        return
    m = re.match(r"^(\w+):$", s)
    if m:
        out.append(local_label(m.group(1), fn) + ":")     # Go labels are function-local
        return
    parts = s.split(None, 1)
    mn = parts[0]
    ops = split_ops(parts[1]) if len(parts) > 1 else []
    if mn == "RET":
        if fn.frame:
            out.append("\taddq $%d, %%rsp" % fn.frame)
        out.append("\tret")
        return
    if mn == "PCALIGN":
        n = int(ops[0][1:], 0)
        out.append("\t.p2align %d" % (n.bit_length() - 1))
        return
    if mn in JCC:
        out.append("\t%s %s" % (JCC[mn], local_label(ops[0], fn)))
        return
    if mn in SSE:
        out.append("\t%s %s" % (SSE[mn], ", ".join(operand(o, "q", fn) for o in ops)))
        return
    if mn in ZX:
        att, s0, s1 = ZX[mn]
        out.append("\t%s %s, %s" % (att, operand(ops[0], s0, fn), operand(ops[1], s1, fn)))
        return
    m = re.fullmatch(r"CMOV([QL])(LT|GE|LE|EQ)", mn)
    if m:
        sz = m.group(1).lower()
        out.append("\tcmov%s %s, %s" % (CMOV_CC[m.group(2)], operand(ops[0], sz, fn), operand(ops[1], sz, fn)))
        return
    if mn in SIMPLE:
        att, sz = SIMPLE[mn]
        t = [operand(o, sz, fn) for o in ops]
        if mn.startswith("CMP"):
            t = t[::-1]          # Plan 9: CMP a, b -> flags of a - b; AT&T: cmp b, a
        if mn.startswith(("SHL", "SHR", "SAR")) and len(ops) == 2 and not ops[0].startswith("$"):
            t[0] = "%cl"
        if mn == "MOVQ" and ops[0].startswith("$") and ops[1] in GP:
            att = "movabsq" if int(ops[0][1:], 0) > 0x7fffffff else "movq"
        out.append("\t%s %s" % (att, ", ".join(t)))
        return
    raise ValueError("unknown mnemonic %s in %r" % (mn, raw))


def main(src, dst):
    fn = Fn()
    out = ["# generated by oracle/p9_to_gas.py from %s -- do not commit" % src]
    with open(src, encoding="utf-8") as f:
        for ln, line in enumerate(f, 1):
            try:
                translate_line(line, fn, out)
            except ValueError as e:
                raise SystemExit("%s:%d: %s" % (src, ln, e))
    out.append('\t.section .note.GNU-stack,"",@progbits')
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
