//**************************************************************************
// file name: shim_test.cc
//**************************************************************************
// One driver, two builds: against the shim headers (hackrfdiags_b200/shim/include,
// -> shim_test) and against the reference's own headers and sources
// (oracle/Makefile -> oracle/_ref/shim_test_ref).  It uses nothing but the public
// operations both sets of classes share, in the way the reference's callers do
// (IqDataProcessor.cc:991-1034 for the demodulators, am.cc:31-68 and
// BasebandDataProcessor.cc:648-687 for the modulators), so byte-identical output
// files are the drop-in proof.  tests/test_gpu_shim.py runs both.
//
//   shim_test iqdp <none|am|fm|wbfm|lsb|usb> <iq2048k.s8> <pcm.s16> [squelch threshold dBFS]
//       Radio.cc's use of IqDataProcessor: the four demodulators handed in, 262144-byte blocks through
//       acceptIqData (DataConsumer.cc:319-351), signal state / magnitude callbacks printed per block
//   shim_test rx <am|fm|wbfm|lsb|usb> <iq256k.s8> <pcm.s16> [gain]
//   shim_test tx <am|fm|wbfm|lsb|usb> <pcm.s16> <iq.s8> [index-or-deviation]
//**************************************************************************
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "AmDemodulator.h"
#include "FmDemodulator.h"
#include "WbFmDemodulator.h"
#include "SsbDemodulator.h"
#include "AmModulator.h"
#include "FmModulator.h"
#include "WbFmModulator.h"
#include "SsbModulator.h"
#include "IqDataProcessor.h"

// Radio.cc:15,413: the receive gain the squelch refers the signal level to
uint32_t radio_adjustableReceiveGainInDb = 16;

// the classes print through this (diagUi.cc:2881; am.cc:93-110)
void nprintf(FILE *s,const char *formatPtr, ...)
{
  va_list args;
  va_start(args,formatPtr);
  vfprintf(s,formatPtr,args);
  va_end(args);
}

static std::vector<int16_t> pcmSink;
static unsigned callbackCount;

static void pcmCallback(int16_t *bufferPtr,uint32_t bufferLength)
{
  pcmSink.insert(pcmSink.end(),bufferPtr,bufferPtr + bufferLength);
  callbackCount++;
}

static std::vector<char> readFile(const char *namePtr)
{
  std::vector<char> data;
  FILE *f = fopen(namePtr,"rb");
  if (f == NULL) { perror(namePtr); exit(2); }
  char chunk[65536];
  size_t n;
  while ((n = fread(chunk,1,sizeof chunk,f)) > 0) data.insert(data.end(),chunk,chunk + n);
  fclose(f);
  return (data);
}

static void writeFile(const char *namePtr,const void *dataPtr,size_t n)
{
  FILE *f = fopen(namePtr,"wb");
  if (f == NULL) { perror(namePtr); exit(2); }
  fwrite(dataPtr,1,n,f);
  fclose(f);
}

template <class Demodulator>
static void runDemodulator(Demodulator &d,std::vector<char> &iq)
{
  // IqDataProcessor hands over 32768 bytes per call; the middle of the run
  // also exercises a reset (resetDemodulator is on the public interface)
  const size_t block = 32768;
  size_t offset = 0;
  unsigned calls = 0;
  while (offset < iq.size())
  {
    size_t n = iq.size() - offset;
    if (n > block) n = block;
    d.acceptIqData((int8_t *)&iq[offset],(uint32_t)n);
    offset += n;
    calls++;
    if (calls == 3) d.resetDemodulator();
  }
  d.displayInternalInformation();
  fprintf(stderr,"calls %u callbacks %u\n",calls,callbackCount);
}

static void signalStateCallback(bool signalPresent,void *contextPtr)
{
  (void)contextPtr;
  fprintf(stderr,"signal %d\n",signalPresent ? 1 : 0);
}

static void signalMagnitudeCallback(uint32_t signalMagnitude,void *contextPtr)
{
  (void)contextPtr;
  fprintf(stderr,"magnitude %u\n",(unsigned)signalMagnitude);
}

// What Radio.cc does with an IqDataProcessor (Radio.cc:150-200, 2396-2633) and what dataConsumerThread
// feeds it (DataConsumer.cc:319-351)
static void runIqDataProcessor(const char *mode,std::vector<char> &iq,bool haveThreshold,int threshold)
{
  char host[] = "127.0.0.1";
  IqDataProcessor processor(host,8001);
  AmDemodulator am(pcmCallback);
  FmDemodulator fm(pcmCallback);
  WbFmDemodulator wbFm(pcmCallback);
  SsbDemodulator ssb(pcmCallback);
  processor.setAmDemodulator(&am);
  processor.setFmDemodulator(&fm);
  processor.setWbFmDemodulator(&wbFm);
  processor.setSsbDemodulator(&ssb);
  IqDataProcessor::demodulatorType type = IqDataProcessor::None;
  if (strcmp(mode,"am") == 0) type = IqDataProcessor::Am;
  if (strcmp(mode,"fm") == 0) type = IqDataProcessor::Fm;
  if (strcmp(mode,"wbfm") == 0) type = IqDataProcessor::WbFm;
  if (strcmp(mode,"lsb") == 0) type = IqDataProcessor::Lsb;
  if (strcmp(mode,"usb") == 0) type = IqDataProcessor::Usb;
  processor.setDemodulatorMode(type);
  if (haveThreshold) processor.setSignalDetectThreshold(threshold);
  processor.registerSignalStateCallback(signalStateCallback,NULL);
  processor.registerSignalMagnitudeCallback(signalMagnitudeCallback,NULL);
  processor.enableSignalNotification();
  processor.enableSignalMagnitudeNotification();
  const size_t block = 262144;
  size_t offset = 0;
  unsigned calls = 0;
  while (offset < iq.size())
  {
    size_t n = iq.size() - offset;
    if (n > block) n = block;
    processor.acceptIqData(calls,(int8_t *)&iq[offset],n);
    offset += n;
    calls++;
    if (calls == 2) { am.setDemodulatorGain(150.0f); fm.resetDemodulator(); ssb.setDemodulatorGain(450.0f); }
    if (calls == 4) radio_adjustableReceiveGainInDb = 24;
  }
  processor.displayInternalInformation();
  ssb.displayInternalInformation();
  fprintf(stderr,"calls %u callbacks %u\n",calls,callbackCount);
}

template <class Modulator>
static std::vector<int8_t> runModulator(Modulator &m,std::vector<char> &pcmBytes)
{
  // am.cc:31-68: 512 samples in, outputBufferLength bytes out, until the input ends
  std::vector<int8_t> out;
  std::vector<int8_t> block(512 * 512);
  int16_t *pcm = (int16_t *)&pcmBytes[0];
  size_t total = pcmBytes.size() / 2, offset = 0;
  unsigned calls = 0;
  while (offset < total)
  {
    size_t n = total - offset;
    if (n > 512) n = 512;
    uint32_t produced = 0;
    m.acceptData(&pcm[offset],(uint32_t)n,&block[0],&produced);
    out.insert(out.end(),block.begin(),block.begin() + produced);
    offset += n;
    calls++;
    if (calls == 2) m.resetModulator();
  }
  m.displayInternalInformation();
  return (out);
}

int main(int argc,char **argv)
{
  if (argc < 5)
  {
    fprintf(stderr,"usage: %s rx|tx|iqdp am|fm|wbfm|lsb|usb <in> <out> [parameter]\n",argv[0]);
    return (2);
  }
  const char *dir = argv[1], *mode = argv[2];
  std::vector<char> in = readFile(argv[3]);
  const bool haveParameter = argc > 5;
  const float parameter = haveParameter ? (float)atof(argv[5]) : 0;

  if (strcmp(dir,"iqdp") == 0)
  {
    runIqDataProcessor(mode,in,haveParameter,(int)parameter);
    writeFile(argv[4],pcmSink.data(),pcmSink.size() * sizeof(int16_t));
  }
  else if (strcmp(dir,"rx") == 0)
  {
    if (strcmp(mode,"am") == 0)
    {
      AmDemodulator d(pcmCallback);
      if (haveParameter) d.setDemodulatorGain(parameter);
      runDemodulator(d,in);
    }
    else if (strcmp(mode,"fm") == 0)
    {
      FmDemodulator d(pcmCallback);
      if (haveParameter) d.setDemodulatorGain(parameter);
      runDemodulator(d,in);
    }
    else if (strcmp(mode,"wbfm") == 0)
    {
      WbFmDemodulator d(pcmCallback);
      if (haveParameter) d.setDemodulatorGain(parameter);
      runDemodulator(d,in);
    }
    else
    {
      SsbDemodulator d(pcmCallback);
      if (strcmp(mode,"usb") == 0) d.setUsbDemodulationMode(); else d.setLsbDemodulationMode();
      if (haveParameter) d.setDemodulatorGain(parameter);
      runDemodulator(d,in);
    }
    writeFile(argv[4],pcmSink.data(),pcmSink.size() * sizeof(int16_t));
  }
  else
  {
    std::vector<int8_t> out;
    if (strcmp(mode,"am") == 0)
    {
      AmModulator m;
      if (haveParameter) { m.setModulationIndex(parameter); m.setModulationIndex(1.5f); /* rejected */ }
      out = runModulator(m,in);
    }
    else if (strcmp(mode,"fm") == 0)
    {
      FmModulator m;
      if (haveParameter) m.setFrequencyDeviation(parameter);
      out = runModulator(m,in);
    }
    else if (strcmp(mode,"wbfm") == 0)
    {
      WbFmModulator m;
      if (haveParameter) m.setFrequencyDeviation(parameter);
      out = runModulator(m,in);
    }
    else
    {
      SsbModulator m;
      if (strcmp(mode,"usb") == 0) m.setUsbModulationMode(); else m.setLsbModulationMode();
      out = runModulator(m,in);
    }
    writeFile(argv[4],out.data(),out.size());
  }
  return (0);
}
