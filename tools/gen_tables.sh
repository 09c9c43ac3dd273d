#!/bin/sh
# Regenerates vvc_intra_b200/csrc/vvc_rom_tables.h from the reference's ROM (needs /root/reference
# and oracle/_ref/libvtmref.a, i.e. the build container).  The generated header is committed.
set -e
cd "$(dirname "$0")/.."
make -f oracle/Makefile.ref -j8 oracle/_ref/libvtmref.a >/dev/null
SRC=/root/reference/VVC_project/source
g++ -std=c++11 -O1 -w -msse4.1 -include cstdint -include limits -Ioracle/stub -I/usr/include/python3.12 \
    -I$SRC/Lib -I$SRC/Lib/CommonLib -I$SRC/Lib/libmd5 oracle/dump_tables.cpp oracle/_ref/libvtmref.a -pthread -o oracle/_ref/dump_tables
oracle/_ref/dump_tables > vvc_intra_b200/csrc/vvc_rom_tables.h
wc -c vvc_intra_b200/csrc/vvc_rom_tables.h
