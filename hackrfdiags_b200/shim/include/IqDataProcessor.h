//**************************************************************************
// file name: IqDataProcessor.h   (libhrdshim: the reference's class on the libhrd_b200 C ABI)
//**************************************************************************
// Same public interface as radioDiags/hdr_diags/IqDataProcessor.h:21-70, so that Radio.cc and
// DataConsumer.cc compile unchanged.  One object is one stream of an Rx batch at the 2.048 MS/s
// entry: front end, squelch gate and the selected demodulator all run inside hrd_rx_process
// (IqDataProcessor.cc:926-1038); the four demodulator objects handed in through
// set<X>Demodulator() keep their roles as parameter holders (gain, sideband, resets) and as the
// owners of the PCM callbacks.
//**************************************************************************
#ifndef _IQDATAPROCESSOR_H_
#define _IQDATAPROCESSOR_H_

#include <stdint.h>

#include "AmDemodulator.h"
#include "FmDemodulator.h"
#include "WbFmDemodulator.h"
#include "SsbDemodulator.h"

struct HrdShimIqdp; // private: the hrd_batch_t and its staging buffers

class IqDataProcessor
{
  public:

  enum demodulatorType {None=0, Am=1, Fm=2, WbFm = 3, Lsb = 4, Usb = 5};

  IqDataProcessor(char *hostIpAddress,int hostPort);
  ~IqDataProcessor(void);

  void setDemodulatorMode(demodulatorType mode);
  void setAmDemodulator(AmDemodulator *demodulatorPtr);
  void setFmDemodulator(FmDemodulator *demodulatorPtr);
  void setWbFmDemodulator(WbFmDemodulator *demodulatorPtr);
  void setSsbDemodulator(SsbDemodulator *demodulatorPtr);
  void setSignalDetectThreshold(int32_t threshold);

  uint32_t reduceSampleRate(int8_t *bufferPtr,uint32_t bufferLength);
  void downconvertByFsOver4(int8_t *bufferPtr,uint32_t byteCount);
  void upconvertByFsOver4(int8_t *bufferPtr,uint32_t byteCount);

  void acceptIqData(unsigned long timeStamp,
                    int8_t *bufferPtr,
                    unsigned long byteCount);

  void enableSignalNotification(void);
  void disableSignalNotification(void);

  void registerSignalStateCallback(
      void (*signalCallbackPtr)(bool signalPresent,
                                void *contextPtr),
      void *contextPtr);

  void enableSignalMagnitudeNotification(void);
  void disableSignalMagnitudeNotification(void);

  void registerSignalMagnitudeCallback(
      void (*callbackPtr)(uint32_t signalMagnitude,void *contextPtr),
      void *contextPtr);

  void enableIqDump(void);
  void disableIqDump(void);
  bool isIqDumpEnabled(void);

  void displayInternalInformation(void);

  // what reduceSampleRate() produced (the reference keeps it in a private member of the same name)
  int8_t decimatedData[32768];

  private:

  IqDataProcessor(const IqDataProcessor &);
  IqDataProcessor &operator=(const IqDataProcessor &);

  HrdShimIqdp *implPtr;
};

#endif // _IQDATAPROCESSOR_H_
