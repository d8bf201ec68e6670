//**************************************************************************
// file name: WbFmDemodulator.h  (libhrd_b200 shim)
//**************************************************************************
// Drop-in for the reference's WbFmDemodulator (radioDiags/WbFmDemodulator/WbFmDemodulator.h:23-31): same class name, same
// public operations, same callback contract -- one PCM callback, fired
// synchronously inside acceptIqData(), exactly once per call -- but the signal
// processing runs in libhrd_b200.so (CUDA, sm_100a) as a batch of ONE stream.
// Build the host application against this header instead of the reference's
// and link libhrdshim.a + libhrd_b200.so.  There is no CPU
// fallback: the constructor aborts with a message when no B200 is usable.
//**************************************************************************
#ifndef __WBFMDEMODULATOR__
#define __WBFMDEMODULATOR__

#include <stdint.h>

struct HrdShimRx; // private: the hrd_batch_t and its staging buffers

class WbFmDemodulator
{
  public:

  WbFmDemodulator(
    void (*pcmCallbackPtr)(int16_t *bufferPtr,uint32_t bufferLength));

  ~WbFmDemodulator(void);

  void resetDemodulator(void);
  void setDemodulatorGain(float gain);
  void acceptIqData(int8_t *bufferPtr,uint32_t bufferLength);
  void displayInternalInformation(void);

  private:

  // copying would share the device batch: not supported (the reference's
  // classes own raw pointers and are never copied either)
  WbFmDemodulator(const WbFmDemodulator &);
  WbFmDemodulator &operator=(const WbFmDemodulator &);

  friend class IqDataProcessor; // the shim's IqDataProcessor reads gain / sideband / resets from here

  HrdShimRx *implPtr;
};

#endif // __WBFMDEMODULATOR__
