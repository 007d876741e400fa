// main_gpu.cpp -- `PBDServer --mode gpu`: the reference's own process structure
// (CProgram/src/main.cpp:69-98: sockets_init, Shared, listen_and_accept, sim thread, comm_loop) with
// a CudaStepper behind the IStepper seam.  Everything except this file and CudaStepper.cpp is the
// UNMODIFIED reference, compiled in place from /root/reference/CProgram/src/{Net,Server,Sim}.cpp by
// integration/Makefile: the reference's comm_loop decodes MSG_INIT, its sim_thread_fn calls
// step()/pack_positions() and prints the 1 Hz stats line (Sim.cpp:412-417), its send_positions ships
// MSG_POSITIONS -- so this binary is the proof that the adapter drops in behind the stock client.
// (The reference's main.cpp cannot be reused as is: its parse_args rejects any mode but
// serial|parallel, main.cpp:44-52.  The maintainer's patch is the three lines shown in INTEGRATION.md.)
#include "CudaStepper.h"

int main(int argc, char** argv) {
  int port = 7777, device = 0;
  pbd_options opts{};
  opts.struct_size = sizeof(opts);
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--port" && i + 1 < argc) port = std::atoi(argv[++i]);
    else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (a == "--mode" && i + 1 < argc) { if (std::string(argv[++i]) != "gpu") { std::fprintf(stderr, "this build only has --mode gpu\n"); return 1; } }
    else if (a == "--order" && i + 1 < argc) opts.order_mode = std::string(argv[++i]) == "interleaved" ? PBD_ORDER_INTERLEAVED : PBD_ORDER_STRICT;
    else if (a == "--fast") opts.flags |= PBD_FLAG_FAST_ARITH;
    else if (a == "--help" || a == "-h") { std::printf("Usage: %s --port 7777 [--mode gpu] [--device N] [--order strict|interleaved] [--fast]\n", argv[0]); return 0; }
    else { std::fprintf(stderr, "Unknown arg: %s\n", a.c_str()); return 1; }
  }
  std::setvbuf(stdout, nullptr, _IOLBF, 0);   // the reference printf()s without flushing: keep a piped log line-by-line
  sockets_init();
  CudaStepper gpu(device, &opts);
  Shared sh;
  sh.stepper = &gpu;
  std::printf("[PBDServer] Start. mode=gpu device=%d port=%d\n", device, port);
  std::fflush(stdout);
  SOCKET client = listen_and_accept(port);
  if (client == INVALID_SOCKET) { sockets_shutdown(); return 1; }
  std::thread sim(sim_thread_fn, &sh);
  comm_loop(client, &sh);
  if (sim.joinable()) sim.join();
  sock_close(client);
  sockets_shutdown();
  std::printf("[PBDServer] Shutdown. binds=%u ok=%d\n", gpu.binds(), gpu.ok() ? 1 : 0);
  return gpu.ok() ? 0 : 2;
}
