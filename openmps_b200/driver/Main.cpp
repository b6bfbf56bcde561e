// Main.cpp — the reference's driver (src/OpenMps/Main.cpp:277-406) on top of libopenmps_b200: same command line, same XML
// input, same result/particles_%05d.csv and progress lines, no Boost.  SURVEY.md §8f rank 1.
//
// What differs from a recompile of the reference's Main.cpp against the drop-in headers:
//   * the inner loop `while (T() < nextOutputT) ForwardTime()` is one call (Computer::RunUntil -> mps_run_until), so a step
//     costs no per-step host logic in the driver;
//   * output never stalls the GPU: after an interval the particle state is copied once (device -> host, a few ms at 1 M
//     particles) and handed to a writer thread that formats and writes the CSV (~1 s of CPU at 1 M particles) while the
//     next interval is already being computed.  The progress line of an output is printed when its file has been written,
//     in order.
//   * `--check-io FILE [DIR]` parses FILE, echoes what it read and writes DIR/particles_%05d.csv of the INPUT state without
//     touching a GPU (tests of the formats on machines without one).
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>

#include "ComputingCondition.hpp"
#include "Computer.hpp"
#include "DriverIo.hpp"

namespace
{
	// Timer.hpp:1-32
	class Timer final
	{
		std::chrono::time_point<std::chrono::system_clock> begin;
	public:
		void Start() { begin = std::chrono::system_clock::now(); }
		double Time() const
		{
			const auto end = std::chrono::system_clock::now();
			return static_cast<double>(std::chrono::duration_cast<std::chrono::milliseconds>(end - begin).count()) / 1000.0;
		}
	};

	void System(const char* command)
	{
		const auto ret = std::system(command);
		if (ret < 0) throw std::runtime_error("Error!");
	}

	// one pending output: the state to write and what the progress line needs
	struct Output
	{
		std::vector<OpenMps::Particle> particles;
		std::size_t outputCount = 0, iteration = 0;
		double tComputer = 0;
	};

	// formats and writes the CSVs on its own thread, in submission order; at most `depth` snapshots are in flight
	class Writer final
	{
		std::mutex m;
		std::condition_variable cv;
		std::deque<Output> queue;
		bool done = false;
		std::size_t depth;
		std::string directory;
		const Timer& timer;
		std::thread worker;

		void Run()
		{
			for (;;)
			{
				Output out;
				{
					std::unique_lock<std::mutex> lock(m);
					cv.wait(lock, [this] { return done || !queue.empty(); });
					if (queue.empty()) return;
					out = std::move(queue.front());
				}
				const auto count = OpenMps::DriverIo::OutputToCsv(out.particles, out.outputCount, directory);
				const auto t = std::time(nullptr);
				const auto now = *std::localtime(&t);
				std::cout << OpenMps::DriverIo::ProgressLine(out.tComputer, out.iteration, out.outputCount, count, now, timer.Time()) << std::endl;
				{
					std::lock_guard<std::mutex> lock(m);
					queue.pop_front(); // only now: Submit() counts the snapshot being written as in flight
				}
				cv.notify_all();
			}
		}

	public:
		Writer(const std::string& dir, const Timer& t, const std::size_t maxInFlight = 2)
			: depth(maxInFlight), directory(dir), timer(t), worker([this] { Run(); }) {}

		void Submit(Output&& out)
		{
			std::unique_lock<std::mutex> lock(m);
			cv.wait(lock, [this] { return queue.size() < depth; });
			queue.push_back(std::move(out));
			cv.notify_all();
		}
		void Finish()
		{
			{
				std::lock_guard<std::mutex> lock(m);
				done = true;
			}
			cv.notify_all();
			if (worker.joinable()) worker.join();
		}
		~Writer() { Finish(); }
	};

	int CheckIo(const char* filename, const std::string& directory)
	{
		namespace io = OpenMps::DriverIo;
		std::cout << "Input XML file: " << filename << std::endl;
		const auto xml = io::ReadXmlFile(filename);
		auto&& condition = io::LoadCondition(*xml);
		auto&& environment = io::LoadEnvironment(*xml, condition.OutputInterval);
		auto&& particles = io::LoadParticles(*xml);
		std::printf("condition: eps=%.17g startTime=%.17g endTime=%.17g outputInterval=%.17g\n", condition.Eps, condition.StartTime, condition.EndTime,
			condition.OutputInterval);
		std::printf("environment: dim=%zu l_0=%.17g MaxDt=%.17g MaxDx=%.17g R_e=%.17g Rho=%.17g Nu=%.17g NeighborLength=%.17g n0=%.17g\n", OpenMps::DIM,
			environment.L_0, environment.MaxDt, environment.MaxDx, environment.R_e, environment.Rho, environment.Nu, environment.NeighborLength, environment.N0());
		const auto offset = static_cast<std::size_t>(std::ceil(condition.StartTime / condition.OutputInterval));
		const auto count = io::OutputToCsv(particles, offset, directory);
		std::cout << "wrote " << io::CsvFileName(offset, directory) << ", " << count << " particles not disabled" << std::endl;
		return 0;
	}
}

int main(const int argc, const char* const argv[])
{
	namespace io = OpenMps::DriverIo;
	try
	{
		if (argc >= 3 && std::string(argv[1]) == "--check-io") return CheckIo(argv[2], (argc >= 4) ? argv[3] : "result");

		System("mkdir result");

		const auto filename = (argc == 1) ? "../../Benchmark/Sample/Sample.xml" : argv[1];
		std::cout << "Input XML file: " << filename << std::endl;
		auto xml = io::ReadXmlFile(filename);

		auto&& condition = io::LoadCondition(*xml);
		auto&& environment = io::LoadEnvironment(*xml, condition.OutputInterval);
		auto&& particles = io::LoadParticles(*xml);

		xml.reset(nullptr); // drop the text of the input

		// walls stay where they start (Main.cpp:297-315)
		auto initialPosition = std::make_unique<OpenMps::Vector[]>(particles.size());
		std::transform(particles.cbegin(), particles.cend(), initialPosition.get(), [](const auto& particle) { return particle.X(); });
		const auto positionWall = [&initialPosition](auto i, auto, auto) { return initialPosition[i]; };
		const auto positionWallPre = [](auto, auto) {};

		auto computer = OpenMps::CreateComputer(condition.Eps, environment, positionWall, positionWallPre);
		computer.AddParticles(std::move(particles));

		Timer timer;
		timer.Start();
		Writer writer("result", timer);

		const auto outputIterationOffset = static_cast<std::size_t>(std::ceil(condition.StartTime / condition.OutputInterval));
		const auto snapshot = [&computer](const double t, const std::size_t iteration, const std::size_t outputCount)
		{
			Output out;
			out.particles = computer.Particles(); // one device -> host copy; formatting happens on the writer thread
			out.tComputer = t; out.iteration = iteration; out.outputCount = outputCount;
			return out;
		};
		writer.Submit(snapshot(condition.StartTime, 0, outputIterationOffset)); // the initial state

		double nextOutputT = 0;
		std::size_t iteration = 0;
		const auto endCount = static_cast<std::size_t>(std::ceil((condition.EndTime - condition.StartTime) / condition.OutputInterval));
		for (auto outputCount = decltype(endCount){1}; outputCount <= endCount; outputCount++)
		{
			double tComputer = computer.GetEnvironment().T();
			try
			{
				nextOutputT += condition.OutputInterval;
				iteration += computer.RunUntil(nextOutputT); // while (tComputer < nextOutputT) ForwardTime();
				tComputer = computer.GetEnvironment().T();
				writer.Submit(snapshot(tComputer + condition.StartTime, iteration, outputCount + outputIterationOffset));
			}
			catch (const decltype(computer)::Exception& ex)
			{
				writer.Finish(); // pending progress lines first
				iteration += computer.LastRunSteps(); // the steps of this interval that completed before the failing one
				tComputer = computer.GetEnvironment().T();
				std::cout << "!!!!ERROR!!!!" << std::endl
					<< "#" << (outputCount + outputIterationOffset) << ": t=" << tComputer << " (" << iteration << ")" << std::endl
					<< ex.what() << std::endl;
				break;
			}
		}
		writer.Finish();
		std::cout << "finished" << std::endl;
		return 0;
	}
	catch (const std::exception& ex)
	{
		// the reference lets these escape (terminate -> abort with what()); same text, orderly exit code
		std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << ex.what() << std::endl;
		return 134;
	}
}
