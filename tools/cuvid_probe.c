// Probe: can libnvcuvid be dlopen'ed and does NVDEC report H.264/HEVC support on this box?
// Build: gcc -O2 -o tools/cuvid_probe tools/cuvid_probe.c -ldl
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
typedef struct {
  int eCodecType, eChromaFormat; unsigned nBitDepthMinus8; unsigned reserved1[3];
  unsigned char bIsSupported, nNumNVDECs; unsigned short nOutputFormatMask;
  unsigned nMaxWidth, nMaxHeight, nMaxMBCount; unsigned short nMinWidth, nMinHeight;
  unsigned char bIsHistogramSupported, nCounterBitDepth; unsigned short nMaxHistogramBins;
  unsigned reserved3[10];
} CAPS;
typedef int (*cuInit_t)(unsigned); typedef int (*cuDeviceGet_t)(int*, int);
typedef int (*cuCtxCreate_t)(void**, unsigned, int); typedef int (*caps_t)(CAPS*);
typedef int (*retain_t)(void**, int); typedef int (*setcur_t)(void*);
int main(void) {
  const char* names[] = {"libnvcuvid.so.1", "/usr/local/nvidia/lib/libnvcuvid.so.1", "/usr/local/nvidia/lib64/libnvcuvid.so.1",
                         "/usr/lib/x86_64-linux-gnu/libnvcuvid.so.1"};
  void* h = 0; for (int i = 0; i < 4 && !h; i++) { h = dlopen(names[i], RTLD_NOW); printf("dlopen %s -> %s\n", names[i], h ? "OK" : dlerror()); }
  if (!h) return 1;
  void* cu = dlopen("libcuda.so.1", RTLD_NOW); if (!cu) { printf("no libcuda\n"); return 1; }
  cuInit_t cuInit = (cuInit_t)dlsym(cu, "cuInit"); cuDeviceGet_t cuDeviceGet = (cuDeviceGet_t)dlsym(cu, "cuDeviceGet");
  retain_t retain = (retain_t)dlsym(cu, "cuDevicePrimaryCtxRetain"); setcur_t setcur = (setcur_t)dlsym(cu, "cuCtxSetCurrent");
  int dev; void* ctx; int r = cuInit(0); printf("cuInit %d\n", r); cuDeviceGet(&dev, 0); r = retain(&ctx, dev); printf("retain %d\n", r); setcur(ctx);
  caps_t caps = (caps_t)dlsym(h, "cuvidGetDecoderCaps"); if (!caps) { printf("no cuvidGetDecoderCaps\n"); return 1; }
  int codecs[] = {4, 8, 10, 11}; const char* cn[] = {"H264", "HEVC", "VP9", "AV1"};
  for (int i = 0; i < 4; i++) { CAPS c; memset(&c, 0, sizeof c); c.eCodecType = codecs[i]; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0;
    r = caps(&c); printf("%s: rc=%d supported=%d nNVDECs=%d fmtmask=0x%x max=%ux%u maxMB=%u min=%ux%u\n", cn[i], r, c.bIsSupported, c.nNumNVDECs,
                         c.nOutputFormatMask, c.nMaxWidth, c.nMaxHeight, c.nMaxMBCount, c.nMinWidth, c.nMinHeight); }
  const char* syms[] = {"cuvidCreateVideoParser","cuvidParseVideoData","cuvidDestroyVideoParser","cuvidCreateDecoder","cuvidDestroyDecoder","cuvidDecodePicture","cuvidGetDecodeStatus","cuvidMapVideoFrame64","cuvidUnmapVideoFrame64","cuvidCtxLockCreate","cuvidReconfigureDecoder"};
  for (int i = 0; i < 11; i++) printf("sym %s %s\n", syms[i], dlsym(h, syms[i]) ? "ok" : "MISSING");
  return 0;
}
