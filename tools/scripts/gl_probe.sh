#!/bin/bash
# Probes the GPU box for what the reference's own GL path would need:
# Mesa EGL / GLESv2 libraries + headers, cmake, libjpeg.  Output is committed under profiles/ as the evidence for
# "reference not runnable on this box" (bench.py's CPU arm is then the oracle port, kind "port").
echo "== date: $(date -u +%FT%TZ)  host: $(hostname)  nproc: $(nproc)"
echo "== ldconfig -p | grep -E 'libEGL|libGLESv2|libGL\.|libOSMesa|libgbm|libX11|libjpeg'"
ldconfig -p | grep -E 'libEGL|libGLESv2|libGL\.|libOSMesa|libgbm|libX11|libjpeg' || echo "(none)"
echo "== headers"
for d in /usr/include/EGL /usr/include/GLES3 /usr/include/GLES2 /usr/include/GL /usr/include/KHR; do
  if [ -d "$d" ]; then echo "$d: present"; else echo "$d: absent"; fi
done
for f in /usr/include/jpeglib.h; do [ -f "$f" ] && echo "$f: present" || echo "$f: absent"; done
echo "== find / -name 'libEGL*' -o -name 'libGLESv2*' -o -name 'swrast_dri.so' -o -name 'libgallium*' (first 20)"
find / -xdev \( -name 'libEGL*' -o -name 'libGLESv2*' -o -name 'swrast_dri.so' -o -name 'libgallium*' -o -name 'libOSMesa*' \) 2>/dev/null | head -20
echo "== tools"
for t in cmake eglinfo glxinfo Xvfb; do printf "%s: " "$t"; command -v "$t" || echo "absent"; done
echo "== /root/reference present: $([ -d /root/reference ] && echo yes || echo no)"
echo "== nvidia EGL (GPU-accelerated GL is not the CPU baseline BASELINE.md asks for, listed for completeness)"
ls /usr/share/glvnd/egl_vendor.d 2>/dev/null || echo "(no glvnd vendor dir)"
