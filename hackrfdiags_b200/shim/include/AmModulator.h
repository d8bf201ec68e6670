//**************************************************************************
// file name: AmModulator.h  (libhrd_b200 shim)
//**************************************************************************
// Drop-in for the reference's AmModulator (radioDiags/AmModulator/AmModulator.h:22-33): same class name and
// public operations; acceptData() turns bufferLength PCM samples (8000 S/s)
// into bufferLength*512 bytes of interleaved int8 I,Q (2048000 S/s) in the
// caller's buffer.  The work runs in libhrd_b200.so (CUDA, sm_100a) as a
// batch of ONE stream; there is no CPU fallback (the constructor aborts with a
// message when no B200 is usable).  See INTEGRATION.md.
//**************************************************************************
#ifndef __AMMODULATOR__
#define __AMMODULATOR__

#include <stdint.h>

struct HrdShimTx; // private: the hrd_batch_t and its staging buffers

class AmModulator
{
  public:

  AmModulator(void);
  ~AmModulator(void);

  void resetModulator(void);
  void setModulationIndex(float modulationIndex);

  void acceptData(int16_t *bufferPtr,
                  uint32_t bufferLength,
                  int8_t *outputBufferPtr,
                  uint32_t *outputBufferLengthPtr);

  void displayInternalInformation(void);

  private:

  AmModulator(const AmModulator &);
  AmModulator &operator=(const AmModulator &);

  HrdShimTx *implPtr;
};

#endif // __AMMODULATOR__
