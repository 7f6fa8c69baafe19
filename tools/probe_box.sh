#!/bin/bash
# Platform probe run on the GPU box (gpurun): NUMA / PCIe topology as the container sees it, and whether the NVDEC
# user-mode library (libnvcuvid) is reachable from this image.  Output: gpurun_out/probe_box.txt
out=gpurun_out/probe_box.txt
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=index,name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv
echo "== topo"; nvidia-smi topo -m 2>&1 | sed 's/\x1b\[[0-9;]*m//g'
echo "== cpus"; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core|^CPU\(s\)"
echo "== numa nodes"; ls /sys/devices/system/node/ 2>&1 | tr '\n' ' '; echo
for n in /sys/devices/system/node/node*; do echo "$n cpulist=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
echo "== cpuset"; cat /proc/self/status | grep -E "Cpus_allowed_list|Mems_allowed_list"
echo "== gpu pci numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d class=$(cat $d/class) numa=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
echo "== mem"; free -g | head -2
echo "== driver capabilities"; echo "NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
echo "== nvcuvid / nvidia-encode libs"; ls -l /usr/lib/x86_64-linux-gnu/libnvcuvid* /usr/lib/x86_64-linux-gnu/libnvidia-encode* /usr/lib64/libnvcuvid* 2>&1; ldconfig -p | grep -i -E "nvcuvid|nvidia-encode|nvjpeg"
python - <<'PY'
import ctypes
for name in ("libnvcuvid.so.1", "libnvcuvid.so", "libnvidia-encode.so.1", "libnvjpeg.so.12", "libnvjpeg.so"):
    try:
        ctypes.CDLL(name); print("dlopen ok:", name)
    except OSError as e:
        print("dlopen FAILED:", name, "-", e)
try:
    import ctypes as C
    cu = C.CDLL("libcuda.so.1"); nv = C.CDLL("libnvcuvid.so.1")
    cu.cuInit(0)
    dev = C.c_int(); cu.cuDeviceGet(C.byref(dev), 0)
    ctx = C.c_void_p(); cu.cuDevicePrimaryCtxRetain(C.byref(ctx), dev); cu.cuCtxSetCurrent(ctx)
    class CAPS(C.Structure):
        _fields_ = [("eCodecType", C.c_int), ("eChromaFormat", C.c_int), ("nBitDepthMinus8", C.c_uint), ("reserved1", C.c_uint * 3),
                    ("bIsSupported", C.c_ubyte), ("nNumNVDECs", C.c_ubyte), ("nOutputFormatMask", C.c_ushort),
                    ("nMaxWidth", C.c_uint), ("nMaxHeight", C.c_uint), ("nMaxMBCount", C.c_uint),
                    ("nMinWidth", C.c_ushort), ("nMinHeight", C.c_ushort), ("bIsHistogramSupported", C.c_ubyte),
                    ("nCounterBitDepth", C.c_ubyte), ("nMaxHistogramBins", C.c_ushort), ("reserved3", C.c_uint * 10)]
    names = {0: "MPEG1", 1: "MPEG2", 2: "MPEG4", 3: "VC1", 4: "H264", 5: "JPEG", 8: "HEVC", 9: "VP8", 10: "VP9", 11: "AV1"}
    for codec, nm in names.items():
        c = CAPS(); c.eCodecType = codec; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0
        rc = nv.cuvidGetDecoderCaps(C.byref(c))
        print(f"cuvidGetDecoderCaps {nm}: rc={rc} supported={c.bIsSupported} nvdecs={c.nNumNVDECs} max={c.nMaxWidth}x{c.nMaxHeight} outmask={c.nOutputFormatMask}")
except Exception as e:
    print("cuvid caps probe failed:", e)
try:
    import cv2
    print("cv2", cv2.__version__)
    info = cv2.getBuildInformation()
    for line in info.splitlines():
        if any(k in line for k in ("FFMPEG", "avcodec", "avformat", "NVCUVID", "CUDA")):
            print("  ", line.strip())
    for four in ("FFV1", "MJPG", "mp4v", "avc1", "H264", "X264", "hev1", "HEVC", "VP90", "AV01", "MPG2"):
        w = cv2.VideoWriter(f"/tmp/probe_{four}.avi" if four in ("FFV1", "MJPG") else f"/tmp/probe_{four}.mp4", cv2.VideoWriter_fourcc(*four), 30, (320, 240))
        print("  writer", four, w.isOpened()); w.release()
except Exception as e:
    print("cv2 probe failed:", e)
PY
} > $out 2>&1
echo probe done
