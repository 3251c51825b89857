# Host layer: liblpr381.so (the C++ mirror of the reference's controllers) and the lpr381 CLI.
# -ffp-contract=off keeps the few host-side floating point expressions (row sums in IsFeasible,
# ratios for the knapsack header) free of FMA, like the C# they mirror.
CXX ?= g++
CXXFLAGS := -std=c++17 -O2 -ffp-contract=off -fPIC -Wall -Wextra
SRC := lp_parser.cpp simplex_host.cpp bnb_host.cpp c_api.cpp
HDR := lp_model.hpp host_util.hpp dotnet_text.hpp ../../include/lpx.h

all: ../liblpr381.so ../lpr381

../liblpr381.so: $(SRC) $(HDR) ../liblpx.so
	$(CXX) $(CXXFLAGS) -shared -o $@ $(SRC) -L.. -llpx -Wl,-rpath,'$$ORIGIN'

../lpr381: cli.cpp $(HDR) ../liblpr381.so
	$(CXX) $(CXXFLAGS) -o $@ cli.cpp -L.. -llpr381 -llpx -Wl,-rpath,'$$ORIGIN'

clean:
	rm -f ../liblpr381.so ../lpr381
