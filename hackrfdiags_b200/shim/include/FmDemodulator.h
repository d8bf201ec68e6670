//**************************************************************************
// file name: FmDemodulator.h  (libhrd_b200 shim)
//**************************************************************************
// Drop-in for the reference's FmDemodulator (radioDiags/FmDemodulator/FmDemodulator.h:23-31): same class name, same
// public operations, same callback contract -- one PCM callback, fired
// synchronously inside acceptIqData(), exactly once per call -- but the signal
// processing runs in libhrd_b200.so (CUDA, sm_100a) as a batch of ONE stream.
// Build the host application against this header instead of the reference's
// and link libhrdshim.a + libhrd_b200.so.  There is no CPU
// fallback: the constructor aborts with a message when no B200 is usable.
//**************************************************************************
#ifndef __FMDEMODULATOR__
#define __FMDEMODULATOR__

#include <stdint.h>

struct HrdShimRx; // private: the hrd_batch_t and its staging buffers

class FmDemodulator
{
  public:

  FmDemodulator(
    void (*pcmCallbackPtr)(int16_t *bufferPtr,uint32_t bufferLength));

  ~FmDemodulator(void);

  void resetDemodulator(void);
  void setDemodulatorGain(float gain);
  void acceptIqData(int8_t *bufferPtr,uint32_t bufferLength);
  void displayInternalInformation(void);

  private:

  // copying would share the device batch: not supported (the reference's
  // classes own raw pointers and are never copied either)
  FmDemodulator(const FmDemodulator &);
  FmDemodulator &operator=(const FmDemodulator &);

  friend class IqDataProcessor; // the shim's IqDataProcessor reads gain / sideband / resets from here

  HrdShimRx *implPtr;
};

#endif // __FMDEMODULATOR__
