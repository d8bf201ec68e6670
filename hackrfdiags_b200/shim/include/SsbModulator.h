//**************************************************************************
// file name: SsbModulator.h  (libhrd_b200 shim)
//**************************************************************************
// Drop-in for the reference's SsbModulator (radioDiags/SsbModulator/SsbModulator.h:23-35): same class name and
// public operations; acceptData() turns bufferLength PCM samples (8000 S/s)
// into bufferLength*512 bytes of interleaved int8 I,Q (2048000 S/s) in the
// caller's buffer.  The work runs in libhrd_b200.so (CUDA, sm_100a) as a
// batch of ONE stream; there is no CPU fallback (the constructor aborts with a
// message when no B200 is usable).  See INTEGRATION.md.
//**************************************************************************
#ifndef __SSBMODULATOR__
#define __SSBMODULATOR__

#include <stdint.h>

struct HrdShimTx; // private: the hrd_batch_t and its staging buffers

class SsbModulator
{
  public:

  SsbModulator(void);
  ~SsbModulator(void);

  void resetModulator(void);
  void setLsbModulationMode(void);
  void setUsbModulationMode(void);

  void acceptData(int16_t *bufferPtr,
                  uint32_t bufferLength,
                  int8_t *outputBufferPtr,
                  uint32_t *outputBufferLengthPtr);

  void displayInternalInformation(void);

  private:

  SsbModulator(const SsbModulator &);
  SsbModulator &operator=(const SsbModulator &);

  HrdShimTx *implPtr;
};

#endif // __SSBMODULATOR__
